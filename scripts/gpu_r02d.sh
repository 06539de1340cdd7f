#!/bin/bash
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "worst|passed|failed|FAILED|Error" | head -40
for v in "SDFG_TC_SPLIT=0" "A=1"; do
  echo "== bench $v"; env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; i=d['inference']
print('ms/step %.3f img/s %.0f | gemm kernels %.3f ms | inf thumb %.3f ms (chain %.3f) feat %.3f ms (chain %.3f)' % (d['ms_per_step'], d['value'], r['kernel_ms_per_step'], i['thumb_only']['ms_per_pass'], i['thumb_only']['field_chain_ms'], i['with_features']['ms_per_pass'], i['with_features']['field_chain_ms']))"
done
