# fused-head backward: parity tests, then K4w / K4x timing (tools/head_bench.py)
mkdir -p gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head" 2>&1 | tail -3
for B in 32 64; do timeout 200 python tools/head_bench.py --B $B 2>&1 | head -1 | cut -c1-330; done | tee gpurun_out/r2u/head_bench_transposed_drain.txt
