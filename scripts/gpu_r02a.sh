#!/bin/bash
# round 2, call A: the whole GPU suite (new parity tests), kernel-level numbers incl. the reference kernels, ncu of the encoder /
# compositing kernels, a short bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_r02a.log; tail -15 gpurun_out/pytest_r02a.log
timeout 600 python scripts/bench_kernels.py 32 > gpurun_out/bench_kernels_r02a.log 2>&1; tail -c 3000 gpurun_out/bench_kernels_r02a.log
timeout 600 bash scripts/gpu_kernels_ncu.sh r02k
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02a.log 2>&1; tail -c 2500 gpurun_out/bench_r02a.log
