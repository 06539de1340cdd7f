#!/bin/bash
# 1 -> 8 GPU scaling of the bench on one box, as the driver launches it (weak scaling: 32 images per GPU) + the data-parallel timeline at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2; do
  NCCL_DEBUG=${NCCL_DEBUG_LEVEL:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.log 2>&1
  tail -1 gpurun_out/scale_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); i=d.get('inference_256') or {}
print('N=%d ms/step %.3f img/s %.0f e2e %.0f | 256^2: %.2f ms %.0f img/s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], i.get('ms_per_pass', 0), i.get('images_per_s', 0)))" || tail -5 gpurun_out/scale_n$N.log
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.log 2>&1; tail -1 gpurun_out/scale_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=1 ms/step %.3f img/s %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 scripts/prof_step_ddp.py > gpurun_out/ddp_timeline_n8.txt 2>&1; tail -40 gpurun_out/ddp_timeline_n8.txt | cut -c1-160
