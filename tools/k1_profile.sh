#!/bin/bash
# tools/k1_profile.sh -- ncu --set full of K1 (bf16, B = 32, the 3 x 64 KiB ring the automatic rule picks) after the plain run exited 0
mkdir -p gpurun_out/k1p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
export IHPR_CALIBRATE=0
./build/kbench 0 32 1 5 | tail -1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_ring_kernel -s 4 -c 1 -f -o gpurun_out/k1p/prof ./build/kbench 0 32 1 5 > gpurun_out/k1p/ncu.log 2>&1
tail -2 gpurun_out/k1p/ncu.log
python tools/ncu_summary.py gpurun_out/k1p/prof.ncu-rep --src fwd_ring_kernel > gpurun_out/k1p/summary.txt 2>&1
rm -f gpurun_out/k1p/prof.ncu-rep
head -60 gpurun_out/k1p/summary.txt
