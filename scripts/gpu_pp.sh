#!/bin/bash
# ping-pong backward chain on/off (same box, alternating)
mkdir -p gpurun_out
for pp in 0 1 0 1; do
  SDFG_TC_PP=$pp timeout 300 python scripts/prof_step.py > gpurun_out/pp_$pp.log 2>&1
  echo "== pp=$pp"; grep -E "tc_chain|span_us" gpurun_out/pp_$pp.log | cut -c1-60,76-100
done
