#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02j.log 2>&1; tail -1 gpurun_out/bench_r02j.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; i=d['inference']
print('ms/step %.3f img/s %.0f e2e %.0f | gemm %.3f ms frac %.3f exec %.1f | inf thumb %.3f (chain %.3f) feat %.3f (chain %.3f)' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['kernel_ms_per_step'], r['frac'], r['achieved_executed'], i['thumb_only']['ms_per_pass'], i['thumb_only']['field_chain_ms'], i['with_features']['ms_per_pass'], i['with_features']['field_chain_ms']))
print(json.dumps(d['inference_256']))
print(json.dumps(d['cpu_baseline']))"
