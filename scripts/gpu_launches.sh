#!/bin/bash
# launch list only (fast): $1 = bench args, $2 = tag
ARGS=${1:-"--steps 1 --warmup 3 --batch 32 --no-cpu-baseline"}
TAG=${2:-r01b32}
SKIP=${3:-500}
CNT=${4:-300}
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
