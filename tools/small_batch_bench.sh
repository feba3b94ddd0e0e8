# small-batch forward (K1c, thread-block cluster + DSMEM merge) against the persistent ring kernel (IHPR_NO_K1C=1), tools/kbench.cu
mkdir -p gpurun_out/r2k build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/kbench tools/kbench.cu -L"integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" -lihpr_b200 -Xlinker -rpath -Xlinker "$PWD/integral-human-pose-regression-for-3d-human-pose-estimation_b200/lib" 2>&1 | tail -2
for B in 1 2 4 8; do for dt in 0 1; do echo -n "K1c  : "; ./build/kbench 0 $B $dt 100 | tail -1; echo -n "ring : "; IHPR_NO_K1C=1 ./build/kbench 0 $B $dt 100 | tail -1; done; done > gpurun_out/r2k/kbench_small.txt 2>&1
cat gpurun_out/r2k/kbench_small.txt
python tools/host_overhead.py > gpurun_out/r2k/host_overhead.txt 2>&1; cat gpurun_out/r2k/host_overhead.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2k/bench.json').read().strip().splitlines()[-1]); t=d['train']; print(d['value'], d['e2e']['value'], t['value'], t['ms_per_step'], t.get('cuda_graph'))"
