#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder.py -x -q 2>&1 | tail -25
timeout 300 python -m pytest tests/test_gpu_tc.py -q -s -k "reference_fixture" 2>&1 | grep -E "worst|Error|passed|failed" | head
