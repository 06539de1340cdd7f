#!/bin/bash
mkdir -p gpurun_out
env TAG=rbit timeout 300 python scripts/dbg_fullsize.py 1e-4 0 2>&1 | tail -1
env TAG=rbit timeout 300 python scripts/dbg_fullsize.py 1.0 1 2>&1 | tail -1
SDFG_TEST_VAL_TOL=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "reference_fixture" 2>&1 | grep -E "worst|passed|failed"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; i=d['inference']
print('ms/step %.3f img/s %.0f | gemm %.3f ms | inf thumb %.3f (chain %.3f) feat %.3f (chain %.3f) | 256: %.2f ms' % (d['ms_per_step'], d['value'], r['kernel_ms_per_step'], i['thumb_only']['ms_per_pass'], i['thumb_only']['field_chain_ms'], i['with_features']['ms_per_pass'], i['with_features']['field_chain_ms'], d['inference_256']['ms_per_pass']))"
