#!/bin/bash
# tools/k5_sweep3.sh -- K5 with the rotating chunk assignment: parity tests, then split x depth sweep, B = 32 fp32 / bf16, B = 64 fp32, D = 32 / 128 volumes
mkdir -p gpurun_out/k5x build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -q -x 2>&1 | tail -3
export IHPR_CALIBRATE=0
T() { grep FUSED | sed "s/.*fwd+bwd //; s/(1 launch.*//"; }
{
for S in 8 9 10 11 12 13 14 16; do for D in 2 3; do
  echo -n "fp32 B=32 S=$S D=$D : "; IHPR_FUSED_SPLIT=$S IHPR_FUSED_DEPTH=$D ./build/kbench 8 32 0 20 | T
done; done
echo -n "fp32 B=32 default : "; ./build/kbench 8 32 0 20 | T
for S in 3 4 5 6 8; do for D in 2 3; do
  echo -n "bf16 B=32 S=$S D=$D : "; IHPR_FUSED_SPLIT=$S IHPR_FUSED_DEPTH=$D ./build/kbench 8 32 1 20 | T
done; done
echo -n "bf16 B=32 default : "; ./build/kbench 8 32 1 20 | T
echo -n "fp32 B=64 default : "; ./build/kbench 8 64 0 10 | T
echo -n "fp32 B=64 D=32 default : "; ./build/kbench 8 64 0 10 32 | T
echo -n "fp32 B=16 D=128 default : "; ./build/kbench 8 16 0 10 128 | T
} 2>&1 | tee gpurun_out/k5x/sweep3.txt
