#!/bin/bash
mkdir -p gpurun_out
env TAG=rbitc timeout 300 python scripts/dbg_fullsize.py 1e-4 0 2>&1 | tail -1
env TAG=rbitc timeout 300 python scripts/dbg_fullsize.py 1.0 1 2>&1 | tail -1
SDFG_TEST_VAL_TOL=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "reference_fixture" 2>&1 | grep -E "worst|passed|failed"
timeout 300 python scripts/prof_step.py 32 2>&1 | grep -E "span_us|tc_|grid_" | cut -c1-110
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('ms/step %.3f gemm %.3f ms' % (d['ms_per_step'], r['kernel_ms_per_step']))"; done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
