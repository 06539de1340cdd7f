#!/bin/bash
# sweep of the sine split (polynomial pairs per 16-element piece, same value for the training and the inference chain): rebuilt on the box per value
mkdir -p gpurun_out
for v in 2 1 0 3; do
  export SDFG_BUILD_DEFS="-DSDFG_POLY_PAIRS=$v -DSDFG_POLY_PAIRS_TRAIN=$v"
  python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_poly$v.log 2>&1 || { tail -3 gpurun_out/build_poly$v.log; continue; }
  echo "== SDFG_POLY_PAIRS=$v"
  timeout 300 python scripts/prof_step.py 32 2>&1 | grep -E "span_us|fchain" | cut -c1-110
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; i=d['inference']
print('ms/step %.3f gemm %.3f ms | inference chain %.3f / %.3f ms' % (d['ms_per_step'], r['kernel_ms_per_step'], i['thumb_only']['field_chain_ms'], i['with_features']['field_chain_ms']))"
done
