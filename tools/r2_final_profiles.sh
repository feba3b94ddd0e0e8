# round-2 evidence run (1 GPU): full GPU test suite, smoke, default bench, ncu launch list of the bench step, ncu --set full of K4w / K4x / K5
mkdir -p gpurun_out/r2p
python -m pytest tests -m gpu -q > gpurun_out/r2p/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p/pytest.log; tail -3 gpurun_out/r2p/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2p/smoke.log; tail -4 gpurun_out/r2p/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2p/bench_ref.json 2> gpurun_out/r2p/bench_ref.err
python bench.py --steps 2 --warmup 1 --no-train --no-cpu > gpurun_out/r2p/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-train --no-cpu > gpurun_out/r2p/ncu_launches.log 2>&1
python tools/head_bench.py --B 32 --iters 5 > gpurun_out/r2p/head_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"head_bwd_kernel|head_softargmax_kernel" -s 8 -c 6 -o gpurun_out/r2p/prof_head python tools/head_bench.py --B 32 --iters 5 > gpurun_out/r2p/ncu_head.log 2>&1
tail -2 gpurun_out/r2p/ncu_head.log
