#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for mb in 2 64; do
  SDFG_DDP_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_mb$mb.log 2>&1
  tail -1 gpurun_out/bench_n${N}_mb$mb.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); i=d.get('inference_256') or {}
print('N=%d bucket=$mb ms/step %.3f img/s %.0f e2e %.0f | 256^2: %.2f ms %.0f img/s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], i.get('ms_per_pass', 0), i.get('images_per_s', 0)))"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=1 ms/step %.3f' % d['ms_per_step'])"
