#!/bin/bash
mkdir -p gpurun_out
SDFG_TEST_VAL_TOL=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "reference_fixture" 2>&1 | grep -E "worst|passed|failed" 
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s 2>&1 | grep -E "worst gradient|passed|failed|AssertionError" | head -20
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
