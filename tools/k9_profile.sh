# ncu --set full of K9 (after the plain run exited 0), summary pages to gpurun_out/r2v
mkdir -p gpurun_out/r2v
timeout 200 python tools/deconv_bench.py --B 32 --iters 5 > gpurun_out/r2v/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deconv_bn_relu_kernel -s 6 -c 1 -f -o gpurun_out/r2v/prof_k9 python tools/deconv_bench.py --B 32 --iters 5 > gpurun_out/r2v/ncu_k9_full.log 2>&1
tail -2 gpurun_out/r2v/ncu_k9_full.log
ncu -i gpurun_out/r2v/prof_k9.ncu-rep --page raw --csv > gpurun_out/r2v/prof_k9_raw.csv 2>/dev/null
rm -f gpurun_out/r2v/prof_k9.ncu-rep
wc -c gpurun_out/r2v/prof_k9_raw.csv
