#!/bin/bash
# tools/k1_chunk_sweep.sh -- K1 / K2 ring variants at B = 32 (variant 15: 3 x 64 KiB ring, two consumer rounds per barrier hand-shake)
mkdir -p gpurun_out/k1v build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "golden" 2>&1 | tail -2
export IHPR_CALIBRATE=0
{ for dt in 1 0; do for v in 0 11 15 16; do ./build/kbench $v 32 $dt 20 | tail -1; done; done; } 2>&1 | tee gpurun_out/k1v/sweep.txt
