#!/bin/bash
# phase counters of the chain kernels (needs a build with SDFG_BUILD_DEFS=-DSDFG_CHAIN_DEBUG; rebuild without it afterwards):
#   forward chain: CHDBG 4 1000+k lines (SDFG_CHAIN_DBG=1); eikonal pass with the A operand in TMEM: CHDBG 0/2 lines (SDFG_BCHAIN_DBG=1 SDFG_TC_TS=1)
mkdir -p gpurun_out
SDFG_CHAIN_DBG=1 timeout 300 python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline 2>&1 | grep "CHDBG 4 10" > gpurun_out/fchain_phases.log
cat gpurun_out/fchain_phases.log
SDFG_BCHAIN_DBG=1 SDFG_TC_TS=1 timeout 300 python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline 2>&1 | grep -E "CHDBG [02] " > gpurun_out/bchain_phases.log
cat gpurun_out/bchain_phases.log
