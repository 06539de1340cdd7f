#!/bin/bash
# round 2, call B: first run of the folded / collapsed forward chain (tc_fchain.cuh)
mkdir -p gpurun_out
echo "== smoke (tiny fixtures through the new chain)"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | tail -25 > gpurun_out/pytest_tc_r02b.log; tail -12 gpurun_out/pytest_tc_r02b.log
echo "== full size"; timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s 2>&1 | grep -E "worst gradient|passed|failed|Error|assert" | head -20
for v in "SDFG_TC_FWD=old" "SDFG_TC_COLLAPSE=0" "A=1"; do
  echo "== bench $v"; env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; i=d['inference']
print('ms/step %.3f img/s %.0f | gemm kernels %.3f ms | inf thumb %.3f ms (chain %.3f) feat %.3f ms (chain %.3f)' % (d['ms_per_step'], d['value'], r['kernel_ms_per_step'], i['thumb_only']['ms_per_pass'], i['thumb_only']['field_chain_ms'], i['with_features']['ms_per_pass'], i['with_features']['field_chain_ms']))"
done
