#!/bin/bash
# N-GPU bench through torchrun, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "exit $?"; tail -3 gpurun_out/bench_n$N.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "exit $?"; tail -2 gpurun_out/bench_ref_n$N.log | cut -c1-300
