#!/bin/bash
mkdir -p gpurun_out
B='import json,sys
d=json.loads(sys.stdin.readline()); r=d["roofline"]; i=d["inference"]
print("ms/step %.3f img/s %.0f | gemm kernels %.3f ms | inf thumb %.3f ms (chain %.3f) feat %.3f ms (chain %.3f)" % (d["ms_per_step"], d["value"], r["kernel_ms_per_step"], i["thumb_only"]["ms_per_pass"], i["thumb_only"]["field_chain_ms"], i["with_features"]["ms_per_pass"], i["with_features"]["field_chain_ms"]))'
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py tests/test_mesh_path.py -q 2>&1 | tail -6
for v in "A=1" "SDFG_TC_GENERIC_EPI=1" "SDFG_BUILD_DEFS=-DSDFG_WAIT_HINT_NS=0" "SDFG_BUILD_DEFS=-DSDFG_POLY_PAIRS=2" "SDFG_BUILD_DEFS=-DSDFG_POLY_PAIRS=4" "SDFG_BUILD_DEFS=-DSDFG_POLY_PAIRS=0"; do
  echo "== bench $v"; env $v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$B"
done
