#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_render.py -q -x 2>&1 | tail -2
for v in 1 0; do
  export SDFG_EIK_FUSE=$v
  echo "== SDFG_EIK_FUSE=$v"
  timeout 300 python scripts/prof_step.py 32 2>&1 | grep -E "span_us|bwd3|grid_input|grid_backward" | cut -c1-110
  for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('ms/step %.3f gemm %.3f ms e2e %.0f' % (d['ms_per_step'], r['kernel_ms_per_step'], d['e2e']['value']))"; done
done
