#!/bin/bash
# ncu --set full of one kernel inside the training-step bench.  $1 = kernel regex, $2 = tag, $3 = launch skip
KREGEX=${1:-grid_backward_kernel}
TAG=${2:-gridbwd}
SKIP=${3:-3}
ARGS="--steps 1 --warmup 3 --batch 32 --no-cpu-baseline"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 1 -o gpurun_out/prof_$TAG -f python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
