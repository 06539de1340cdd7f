#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export SDFG_EXP_HALFW=1; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$v.log 2>&1
  grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": [0-9.]*' gpurun_out/exp_$v.log
done
