#!/bin/bash
# fused forward chain: parity tests, then forward timing with the chain on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 180 > gpurun_out/pytest_chain.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_chain.log
tail -15 gpurun_out/pytest_chain.log
SDFG_ONLY=tc16 timeout 300 python scripts/bench_field.py 32 > gpurun_out/bench_chain_on.log 2>&1; tail -2 gpurun_out/bench_chain_on.log
SDFG_TC_CHAIN=0 SDFG_ONLY=tc16 timeout 300 python scripts/bench_field.py 32 > gpurun_out/bench_chain_off.log 2>&1; tail -2 gpurun_out/bench_chain_off.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | cut -c1-1200
SDFG_CHAIN_DBG=1 SDFG_ONLY=tc16 timeout 300 python scripts/bench_field.py 8 > gpurun_out/chain_dbg.log 2>&1
grep -c CHDBG gpurun_out/chain_dbg.log
