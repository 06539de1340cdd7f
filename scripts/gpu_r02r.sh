#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_tc.py tests/test_decoder.py -q -x 2>&1 | tail -2
timeout 300 python scripts/bench_kernels.py 32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('composite_forward_feat16','composite_forward_feat32'): print(k, '%.3f ms' % d[k]['ms'], '%.0f GB/s' % d[k]['hbm_GBps'])
print(json.dumps(d.get('inference_forward'))[:200])"
