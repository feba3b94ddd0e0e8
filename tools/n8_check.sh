# 8-GPU check: the default bench line at N=8 (per-rank K5 / K1+K2 times, clocks, measured path choice, train sub-record with the all-reduce cost)
mkdir -p gpurun_out/n8
nvidia-smi topo -m > gpurun_out/n8/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8/bench_n8.json 2> gpurun_out/n8/bench_n8.err; echo "bench rc=$?"
IHPR_CALIBRATE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-train --no-e2e > gpurun_out/n8/bench_n8_nocal.json 2> gpurun_out/n8/bench_n8_nocal.err; echo "bench nocal rc=$?"
tail -2 gpurun_out/n8/bench_n8.err
