#!/bin/bash
# N-GPU bench (N = $1): early table exchange on / off
N=${1:-2}
mkdir -p gpurun_out
for e in 1 0; do
  SDFG_EARLY_EXCHANGE=$e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$e bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_e$e.log 2>&1
  tail -1 gpurun_out/bench_n${N}_e$e.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); i=d.get('inference_256') or {}
print('N=%d early=$e ms/step %.3f img/s %.0f e2e %.0f | 256^2: %.2f ms %.0f img/s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], i.get('ms_per_pass', 0), i.get('images_per_s', 0)))"
done
