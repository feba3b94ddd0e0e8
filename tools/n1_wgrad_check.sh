#!/bin/bash
# tools/n1_wgrad_check.sh -- K11 (deconv weight gradient on tcgen05): tests, timings per cluster / stage-size variant and against the library arm, launch lists
mkdir -p gpurun_out/n1w
timeout 900 python -m pytest tests/test_gpu_deconv_train.py -q -x > gpurun_out/n1w/tests.log 2>&1; echo "train tests rc=$?" | tee -a gpurun_out/n1w/tests.log
tail -3 gpurun_out/n1w/tests.log
for v in ${K11_VARIANTS:-21 22 24 31 32 34}; do timeout 300 python tools/deconv_train_bench.py --B 32 --variant $v; done 2>&1 | tee gpurun_out/n1w/bench.txt
IHPR_DECONV_WGRAD=library timeout 300 python tools/deconv_train_bench.py --B 32 2>&1 | tee -a gpurun_out/n1w/bench.txt
timeout 300 python tools/deconv_train_bench.py --B 64 2>&1 | tee -a gpurun_out/n1w/bench.txt
for v in ${K11_VARIANTS:-21 22 24 31 32 34}; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n1w/launches_$v.csv python tools/deconv_train_bench.py --B 32 --profile --variant $v > gpurun_out/n1w/ncu.log 2>&1
grep -v "^==" gpurun_out/n1w/launches_$v.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if 'k11' in r['Kernel Name']: print($v, r['Kernel Name'][:70], r['Metric Value'])"
done
