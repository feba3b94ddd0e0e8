#!/bin/bash
# tools/n1_profile.sh -- tests, then ncu --set full of the training-mode kernels of row N1 (one forward + backward of the block at B = 32)
mkdir -p gpurun_out/n1
timeout 900 python -m pytest tests/test_gpu_deconv_train.py -q > gpurun_out/n1/tests.log 2>&1; echo "train tests rc=$?" | tee -a gpurun_out/n1/tests.log
tail -4 gpurun_out/n1/tests.log
timeout 200 python tools/deconv_train_bench.py --B 32 | tee gpurun_out/n1/bench.txt || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deconv_bn_relu_kernel|bn_relu|finalize" -f -o gpurun_out/n1/prof python tools/deconv_train_bench.py --B 32 --profile > gpurun_out/n1/ncu.log 2>&1
tail -2 gpurun_out/n1/ncu.log
python tools/ncu_summary.py gpurun_out/n1/prof.ncu-rep > gpurun_out/n1/summary.txt 2>&1
ncu -i gpurun_out/n1/prof.ncu-rep --page raw --csv > gpurun_out/n1/raw.csv 2>/dev/null
rm -f gpurun_out/n1/prof.ncu-rep
cat gpurun_out/n1/summary.txt | head -150
