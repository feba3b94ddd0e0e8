#!/bin/bash
# tools/k11_profile.sh -- ncu --set full of K11 (deconv weight gradient) after the plain run exited 0
mkdir -p gpurun_out/k11
timeout 200 python tools/deconv_train_bench.py --B 32 --profile || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deconv_wgrad_kernel|wgrad_reduce" -f -o gpurun_out/k11/prof python tools/deconv_train_bench.py --B 32 --profile > gpurun_out/k11/ncu.log 2>&1
tail -2 gpurun_out/k11/ncu.log
python tools/ncu_summary.py gpurun_out/k11/prof.ncu-rep > gpurun_out/k11/summary.txt 2>&1
ncu -i gpurun_out/k11/prof.ncu-rep --page raw --csv > gpurun_out/k11/raw.csv 2>/dev/null
rm -f gpurun_out/k11/prof.ncu-rep
cat gpurun_out/k11/summary.txt
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/k11/raw.csv')))
h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:60])
    for k in h:
        if any(x in k for x in ('lts__t_bytes.sum','lts__throughput','l1tex__m_xbar2l1tex_read_bytes.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__inst_executed_pipe_uniform','smsp__cycles_active.avg','lts__d_sectors_fill','dram__throughput')):
            print('   ',k,r[h.index(k)])
PY
