#!/bin/bash
mkdir -p gpurun_out
for v in 0 4 5; do
  SDFG_EXP=$v python scripts/prof_step.py 32 > gpurun_out/exp_$v.log 2>&1
  echo "exp=$v"; grep "tc_chain_fwd" gpurun_out/exp_$v.log | tail -1
done
