#!/bin/bash
# ncu evidence for the current hot path (one GPU).  $1 = extra bench args, $2 = kernel regex for the full capture, $3 = tag
ARGS=${1:-"--steps 1 --warmup 3 --batch 8 --no-cpu-baseline"}
KREGEX=${2:-gemm_f32_kernel}
TAG=${3:-r01}
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
python bench.py $ARGS > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 8 -c 3 -o gpurun_out/prof_$TAG -f python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
tail -2 gpurun_out/plain_$TAG.log | cut -c1-400
