#!/bin/bash
# ncu --set full of the encoder / compositing kernels (and the reference's kernels from oracle/_ref) at N = 3.1 M real ray samples.
# $1 = tag.  One GPU, one ncu invocation.
TAG=${1:-r02k}
mkdir -p gpurun_out
python scripts/run_kernels_once.py 32 2 > gpurun_out/kernels_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"sample_rays|grid_forward|grid_backward|grid_input_backward|sh_forward|composite_forward|composite_backward|kernel_grid|kernel_sh" \
    -s 14 -c 14 -o gpurun_out/prof_$TAG -f python scripts/run_kernels_once.py 32 2 > gpurun_out/ncu_kernels_$TAG.log 2>&1
echo "kernels ncu exit $?"
