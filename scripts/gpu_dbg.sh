#!/bin/bash
mkdir -p gpurun_out
SDFG_CHAIN_DBG=1 SDFG_ONLY=tc16 timeout 300 python scripts/bench_field.py 8 > gpurun_out/chain_dbg.log 2>&1
grep -c CHDBG gpurun_out/chain_dbg.log; tail -1 gpurun_out/chain_dbg.log
