# release build: the whole GPU suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2f/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r2f/smoke.txt
timeout 900 python bench.py > gpurun_out/r2f/bench_default.json 2> gpurun_out/r2f/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2f/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f/bench_reference.json 2> gpurun_out/r2f/bench_reference.err; echo "reference rc=$?"
