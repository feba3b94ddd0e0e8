#!/bin/bash
# tools/k1c_sweep.sh -- K1c (cluster forward, register-prefetched) per cluster size against the persistent ring kernel; also B = 32 (the ring kernel must be unchanged by the consumer refactor)
mkdir -p gpurun_out/k1c build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "small_batch or golden or cluster" 2>&1 | tail -3
{
for B in 1 2 3 4 6 8; do for dt in 0 1; do
  for cs in 0 2 4 8; do echo -n "B=$B dt=$dt cs=$cs : "; IHPR_K1C_CS=$cs ./build/kbench 0 $B $dt 100 | tail -1 | sed 's/fwd+bwd.*//'; done
  echo -n "B=$B dt=$dt auto : "; ./build/kbench 0 $B $dt 100 | tail -1 | sed 's/fwd+bwd.*//'
done; done
for dt in 0 1; do echo -n "B=32 dt=$dt ring : "; ./build/kbench 0 32 $dt 20 | tail -1; done
} 2>&1 | tee gpurun_out/k1c/sweep.txt
