#!/bin/bash
# A/B of the rounding-bit variants (SDFG_RBIT = 1 / 0): rebuilt on the box, per-kernel times of the training step + bench line
mkdir -p gpurun_out
for v in 1 0; do
  export SDFG_BUILD_DEFS="-DSDFG_RBIT=$v"
  python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_rbit$v.log 2>&1 || { tail -3 gpurun_out/build_rbit$v.log; continue; }
  echo "== SDFG_RBIT=$v"
  timeout 300 python scripts/prof_step.py 32 2>&1 | grep -E "span_us|tc_|grid_" | cut -c1-110
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('ms/step %.3f gemm %.3f ms' % (d['ms_per_step'], r['kernel_ms_per_step']))"
  [ $v = 1 ] && env TAG=rbit timeout 300 python scripts/dbg_fullsize.py 1e-4 0 2>&1 | tail -1
done
