#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the bench line (default flags) and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; q=d['inference_256']
print('ms/step %.3f img/s %.0f e2e %.0f launches %s | gemm %.3f ms frac %.3f hbm_frac %.3f | 256: %.2f ms %.0f img/s graphed %.0f | cpu %s | clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r['kernel_ms_per_step'], r['frac'], r['hbm_frac'], q['ms_per_pass'], q['images_per_s'], q['graphed_images_per_s'], json.dumps(d['cpu_baseline'])[:160], json.dumps(d['clocks'])))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
