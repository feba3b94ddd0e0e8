#!/bin/bash
# tools/k5_sweep4.sh -- K5 on 64 KiB chunks (3-stage ring), IHPR_FUSED_CHUNK=64: parity of one configuration against K1 + K2 (kbench prints it), then timings
mkdir -p gpurun_out/k5x build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
export IHPR_CALIBRATE=0 IHPR_FUSED_CHUNK=64
T() { grep -E "FUSED|one-launch vs" | sed "s/.*fwd+bwd //; s/(1 launch.*//; s/.*one-launch vs K1+K2: //" | tr '\n' ' '; echo; }
{
for S in 8 16; do for D in 1 2 3; do
  echo -n "fp32 B=32 chunk64 S=$S D=$D : "; IHPR_FUSED_SPLIT=$S IHPR_FUSED_DEPTH=$D ./build/kbench 8 32 0 20 | T
done; done
for S in 2 4 8; do for D in 1 2 3; do
  echo -n "bf16 B=32 chunk64 S=$S D=$D : "; IHPR_FUSED_SPLIT=$S IHPR_FUSED_DEPTH=$D ./build/kbench 8 32 1 20 | T
done; done
} 2>&1 | tee gpurun_out/k5x/sweep4.txt
