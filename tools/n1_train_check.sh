#!/bin/bash
# tools/n1_train_check.sh -- GPU check of the training deconv block (row N1): parity tests, K9 inference regression, timings, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deconv_train.py -x -q > gpurun_out/n1_tests.log 2>&1; echo "train tests rc=$?" | tee -a gpurun_out/n1_tests.log
tail -15 gpurun_out/n1_tests.log
timeout 600 python -m pytest tests/test_gpu_deconv.py -x -q > gpurun_out/n1_k9_tests.log 2>&1; echo "k9 tests rc=$?" | tee -a gpurun_out/n1_k9_tests.log
tail -3 gpurun_out/n1_k9_tests.log
for B in 32 64; do timeout 300 python tools/deconv_train_bench.py --B $B; done 2>&1 | tee gpurun_out/n1_train_bench.txt
timeout 300 python tools/deconv_bench.py --B 32 2>&1 | tee -a gpurun_out/n1_train_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n1_launches.csv python tools/deconv_train_bench.py --B 32 --profile > gpurun_out/n1_ncu.log 2>&1
grep -v "^==" gpurun_out/n1_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:90], r['Metric Value'])" | tail -30
