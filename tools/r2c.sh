mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fused_head or deferred or trainer_cuda_graph" > gpurun_out/r2c/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c/pytest.log
grep -E "passed|failed|rc=|^E  " gpurun_out/r2c/pytest.log | head -20
for B in 32 64; do timeout 300 python tools/head_bench.py --B $B > gpurun_out/r2c/head_bench_B$B.txt 2>&1; done
head -1 gpurun_out/r2c/head_bench_B32.txt; head -1 gpurun_out/r2c/head_bench_B64.txt
