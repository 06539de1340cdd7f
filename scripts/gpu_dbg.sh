#!/bin/bash
# phase sums of the eikonal chain (needs a build with SDFG_BUILD_DEFS=-DSDFG_CHAIN_DEBUG); forward-chain event log: SDFG_CHAIN_DBG=1
mkdir -p gpurun_out
SDFG_BCHAIN_DBG=1 timeout 300 python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bchain_dbg.log 2>&1
grep CHDBG gpurun_out/bchain_dbg.log
