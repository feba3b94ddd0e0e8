# K9 (deconv + BN + ReLU): parity tests under a short timeout (a protocol bug would hang), timing per cluster size
mkdir -p gpurun_out/r2v
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_deconv.py -m gpu -q -x 2>&1 | tail -3; done
for v in 0 22; do for B in 4 32 64; do timeout 120 python tools/deconv_bench.py --B $B --variant $v 2>&1 | tail -1 | cut -c1-400; done; done | tee gpurun_out/r2v/deconv_bench.txt
