#!/bin/bash
# ncu evidence, end of round 2: launch list of the last training step, --set full of the tensor-core chain kernels, launch list + --set full of
# the decoder kernels of one configs[2] pass.  $1 = tag.  Every ncu run follows a plain run of the same command that exited 0.
TAG=${1:-r02h}
mkdir -p gpurun_out
python scripts/prof_step_once.py train > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_$TAG.csv python scripts/prof_step_once.py train > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list (train) exit $?"
ncu --set full --clock-control none --import-source on -k regex:"tc_fchain|tc_chain_bwd|tc_wgrad" -s 28 -c 7 -o gpurun_out/prof_$TAG -f python scripts/prof_step_once.py train > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture (train) exit $?"
PROF_B=64 python scripts/prof_step_once.py infer256 > gpurun_out/plain3_$TAG.log 2>&1 || { tail -5 gpurun_out/plain3_$TAG.log; exit 1; }
PROF_B=64 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_${TAG}dec.csv python scripts/prof_step_once.py infer256 > gpurun_out/ncu_launchdec_$TAG.log 2>&1
echo "launch list (infer256) exit $?"
# the decoder's kernels of the last pass: 5 convolutions + 2 x 3 transposed-convolution launches + 4 ToRGB = 12 tc_conv_kernel launches (1 + 1 + 2 x (3 + 1 + 1)), 2 blur
PROF_B=64 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_kernel|upconv_blur" -s 28 -c 14 -o gpurun_out/prof_${TAG}dec -f python scripts/prof_step_once.py infer256 > gpurun_out/ncu_fulldec_$TAG.log 2>&1
echo "full capture (infer256) exit $?"
