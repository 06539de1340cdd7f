#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  export SDFG_SCATTER_OVERLAP=$v
  echo "== SDFG_SCATTER_OVERLAP=$v"
  PROF_TIMELINE=1 timeout 300 python scripts/prof_step.py 32 2>&1 | grep -E "span_us|^TL" | grep -E "span|wgrad|grid_backward|head_wgrad|bwd2" | cut -c1-120
  for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('ms/step %.3f gemm %.3f ms e2e %.0f' % (d['ms_per_step'], r['kernel_ms_per_step'], d['e2e']['value']))"; done
done
