#!/bin/bash
# ncu evidence for the current hot path (one GPU): launch list of one bench step + --set full of the three chain kernels.
# $1 = tag
TAG=${1:-r01}
ARGS="--steps 1 --warmup 3 --batch 32 --no-cpu-baseline"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 260 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
python bench.py $ARGS > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_chain|tc_wgrad" -s 8 -c 8 -o gpurun_out/prof_$TAG -f python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
