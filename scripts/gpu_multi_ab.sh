#!/bin/bash
# N-GPU bench: early table-gradient exchange on/off (same box)
N=${1:-2}
mkdir -p gpurun_out
for e in 1 0 1 0; do
SDFG_EARLY_EXCHANGE=$e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$e bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_e$e.log 2>&1
echo "early=$e exit $?"; tail -1 gpurun_out/bench_n${N}_e$e.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2
done
tail -5 gpurun_out/bench_n${N}_e1.log | cut -c1-300
