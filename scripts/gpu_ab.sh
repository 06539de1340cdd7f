#!/bin/bash
# A/B two builds of the library on the same box: lib/libsdfg_A.so vs lib/libsdfg_B.so (alternating, 2 rounds), per-kernel times
mkdir -p gpurun_out
L=sdface-gan_b200/lib
for round in 1 2; do
for v in A B; do
  cp $L/libsdfg_$v.so $L/libsdfg.so
  timeout 300 python scripts/prof_step.py > gpurun_out/ab_${v}${round}.log 2>&1
  echo "== $v$round"; grep -E "tc_chain|span_us" gpurun_out/ab_${v}${round}.log | cut -c1-60,76-100
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*\|"field_chain_ms": [0-9.]*'
done
done
