# 8-GPU check: the default bench line at N=8 exactly as the driver launches it (path metric, per-rank records, train + train.cuda_graph)
mkdir -p gpurun_out/n8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8/bench_n8b.json 2> gpurun_out/n8/bench_n8b.err; echo "bench rc=$?"
tail -2 gpurun_out/n8/bench_n8b.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/n8/bench_n8b.json').read().strip().splitlines()[-1]); t=d['train']; print(d['value'], d['ms_per_step'], t['value'], t['ms_per_step'], t['allreduce']['exposed_ms'], t.get('cuda_graph'))"
