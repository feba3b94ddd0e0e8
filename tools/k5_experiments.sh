#!/bin/bash
# tools/k5_experiments.sh -- K5 (one-launch forward + backward): what the exchange and the L2 policies cost.  Uses a library built with
# -DIHPR_TIMING_EXPERIMENTS (build/lib_exp: tools/build_experiment_lib.sh, run in the container); IHPR_DEBUG_NOXCHG bits: 1 = no exchange (WRONG results), 2 = pass 1 without
# evict_last, 4 = pass 2 without evict_first.  Times from tools/kbench.cu (20 launches), DRAM bytes from ncu.
mkdir -p gpurun_out/k5x build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
export LD_LIBRARY_PATH=$PWD/build/lib_exp:$LD_LIBRARY_PATH
export IHPR_CALIBRATE=0
{
for cfg in "12 3" "12 2" "8 2" "16 3"; do set -- $cfg
  for bits in 0 1 2 4 6; do
    echo -n "S=$1 D=$2 bits=$bits : "; IHPR_FUSED_SPLIT=$1 IHPR_FUSED_DEPTH=$2 IHPR_DEBUG_NOXCHG=$bits ./build/kbench 8 32 0 20 | grep FUSED | sed "s/.*FUSED/FUSED/"
  done
done
} 2>&1 | tee gpurun_out/k5x/times.txt
for cfg in "12 3 0" "12 3 2" "12 3 6" "12 2 0" "8 2 0"; do set -- $cfg
  IHPR_FUSED_SPLIT=$1 IHPR_FUSED_DEPTH=$2 IHPR_DEBUG_NOXCHG=$3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_ring -c 3 --csv --log-file gpurun_out/k5x/ncu_$1_$2_$3.csv env KB_ONLY_FUSED=1 ./build/kbench 8 32 0 3 > /dev/null 2>&1
  echo "S=$1 D=$2 bits=$3: $(grep -v '^==' gpurun_out/k5x/ncu_$1_$2_$3.csv | python -c "
import csv,sys
rows=list(csv.DictReader(sys.stdin))
last=[r for r in rows if r['ID']==rows[-1]['ID']]
print(', '.join(r['Metric Name']+'='+r['Metric Value'] for r in last))")" | tee -a gpurun_out/k5x/dram.txt
done
