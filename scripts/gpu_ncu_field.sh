#!/bin/bash
# ncu --set full of one kernel of the forward-only field bench.  $1 = kernel regex, $2 = tag, $3 = batch
KREGEX=${1:-tc_chain_fwd_kernel}
TAG=${2:-chain}
B=${3:-16}
mkdir -p gpurun_out
SDFG_ONLY=tc16 timeout 300 python scripts/bench_field.py $B > gpurun_out/field_plain_$TAG.log 2>&1 || { tail -5 gpurun_out/field_plain_$TAG.log; exit 1; }
tail -1 gpurun_out/field_plain_$TAG.log
SDFG_ONLY=tc16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/prof_$TAG -f python scripts/bench_field.py $B > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
