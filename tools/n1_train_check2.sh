#!/bin/bash
# tools/n1_train_check2.sh -- second GPU check of row N1 (training): tests, block timings, launch list, whole training step with / without the fused deconv blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deconv_train.py -q > gpurun_out/n1_tests.log 2>&1; echo "train tests rc=$?" | tee -a gpurun_out/n1_tests.log
tail -25 gpurun_out/n1_tests.log
for B in 32 64; do timeout 300 python tools/deconv_train_bench.py --B $B; done 2>&1 | tee gpurun_out/n1_train_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n1_launches.csv python tools/deconv_train_bench.py --B 32 --profile > gpurun_out/n1_ncu.log 2>&1
grep -v "^==" gpurun_out/n1_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if 'cutlass' in r['Kernel Name'] and float(r['Metric Value'])>0: pass
    print(r['Kernel Name'][:70], r['Metric Value'])" | grep -v cutlass | tail -12
for extra in "" "--stock-deconv"; do
  for g in "" "--cuda-graph"; do
    timeout 600 python bench.py --workload train --fused-head $extra $g --steps 20 --warmup 5 2>gpurun_out/train_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', '$extra', '$g', round(d['value'],1), 'samples/s', round(d['ms_per_step'],3), 'ms')" | tee -a gpurun_out/n1_train_step.txt
  done
done
tail -5 gpurun_out/train_err.log
