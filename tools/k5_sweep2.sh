#!/bin/bash
# tools/k5_sweep2.sh -- K5 after the single-trip polling: split / depth / ring-depth sweep (release library), B = 32 fp32 and bf16
mkdir -p gpurun_out/k5x build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
export IHPR_CALIBRATE=0
{
for S in 8 10 12 14 16; do for D in 1 2 3; do for ST in 3 4 6; do
  echo -n "fp32 S=$S D=$D ST=$ST : "; IHPR_FUSED_SPLIT=$S IHPR_FUSED_DEPTH=$D IHPR_FUSED_STAGES=$ST KB_FUSED_ONLY_TIME=1 ./build/kbench 8 32 0 20 | grep FUSED | sed "s/.*fwd+bwd //; s/(1 launch.*//"
done; done; done
for S in 4 6 8; do for D in 1 2 3; do for ST in 3 6; do
  echo -n "bf16 S=$S D=$D ST=$ST : "; IHPR_FUSED_SPLIT=$S IHPR_FUSED_DEPTH=$D IHPR_FUSED_STAGES=$ST ./build/kbench 8 32 1 20 | grep FUSED | sed "s/.*fwd+bwd //; s/(1 launch.*//"
done; done; done
} 2>&1 | tee gpurun_out/k5x/sweep2.txt
