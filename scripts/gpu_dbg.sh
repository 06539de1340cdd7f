#!/bin/bash
mkdir -p gpurun_out
SDFG_CHAIN_DBG=1 timeout 300 python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/fchain_dbg.log 2>&1
grep -c CHDBG gpurun_out/fchain_dbg.log
