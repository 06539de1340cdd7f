#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decoder.py -q -x 2>&1 | tail -12
timeout 300 python scripts/bench_decoder_ops.py 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02q.log 2>&1; tail -1 gpurun_out/bench_r02q.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; i=d['inference']; q=d['inference_256']
print('ms/step %.3f img/s %.0f e2e %.0f | gemm %.3f ms frac %.3f | inf thumb %.3f feat %.3f | 256: %.2f ms (renderer %.2f decoder %.2f) %.0f img/s; graphed %.2f ms %.0f img/s' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['kernel_ms_per_step'], r['frac'], i['thumb_only']['ms_per_pass'], i['with_features']['ms_per_pass'], q['ms_per_pass'], q['renderer_ms'], q['decoder_ms'], q['images_per_s'], q['graphed_ms_per_pass'], q['graphed_images_per_s']))" || tail -20 gpurun_out/bench_r02q.log
