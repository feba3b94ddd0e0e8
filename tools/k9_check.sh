# K9 (deconv3 + BN + ReLU): parity tests under a short timeout (a protocol bug would hang), timing per cluster size, ncu metrics of the kernel
mkdir -p gpurun_out/r2v
timeout 300 python -m pytest tests/test_gpu_deconv.py -m gpu -q -x 2>&1 | tail -5
for v in 0 22; do for B in 4 32 64; do timeout 120 python tools/deconv_bench.py --B $B --variant $v 2>&1 | tail -1 | cut -c1-400; done; done | tee gpurun_out/r2v/deconv_bench.txt
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:deconv_bn -c 8 --csv --log-file gpurun_out/r2v/ncu_k9.csv python tools/deconv_bench.py --B 32 --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2v/ncu_k9.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
for r in rows[1:]:
    if int(r[ii])>=7: print(r[ii], r[ki][:40], r[mi], r[vi])
PY
