#!/bin/bash
# A/B two builds of the library on the same box: lib/libsdfg_A.so vs lib/libsdfg_B.so (alternating, 2 rounds); extra env per run in $ABENV
mkdir -p gpurun_out
L=sdface-gan_b200/lib
for round in 1 2; do
for v in A B; do
 for ov in 0 1; do
  cp $L/libsdfg_$v.so $L/libsdfg.so
  echo "== $v$round overlap=$ov"
  SDFG_OVERLAP=$ov timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
 done
done
done
