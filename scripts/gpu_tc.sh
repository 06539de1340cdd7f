#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
