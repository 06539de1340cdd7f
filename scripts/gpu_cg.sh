#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_cg.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_cg.log
tail -6 gpurun_out/pytest_cg.log
for v in 2 1; do
  SDFG_TC_CG=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/cg_$v.log 2>&1
  echo "cg=$v exit $?"; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": [0-9.]*' gpurun_out/cg_$v.log; tail -2 gpurun_out/cg_$v.log | cut -c1-300
done
