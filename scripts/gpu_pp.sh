#!/bin/bash
# backward-chain variants on the same box (SDFG_TC_TS: A operand in TMEM for the eikonal pass; SDFG_TC_PP: two tiles in flight 0 never / 1 eikonal / 2 always)
mkdir -p gpurun_out
for v in "SDFG_TC_TS=0" "SDFG_TC_TS=1" "SDFG_TC_TS=0 SDFG_TC_PP=2" "SDFG_TC_TS=0" "SDFG_TC_TS=1" "SDFG_TC_TS=0 SDFG_TC_PP=2"; do
  env $v timeout 300 python scripts/prof_step.py > gpurun_out/pp.log 2>&1
  echo "== $v"; grep -E "tc_chain_bwd|span_us" gpurun_out/pp.log | cut -c1-60,76-100
done
