#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoders.py tests/test_gpu_render.py -m gpu -q -x --timeout 180 > gpurun_out/pytest_grid.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_grid.log
tail -5 gpurun_out/pytest_grid.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_quick.log
bash scripts/gpu_launches.sh "--steps 1 --warmup 3 --batch 32 --no-cpu-baseline" r01d 450 260
