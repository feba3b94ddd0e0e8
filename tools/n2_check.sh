# 2-GPU validation: NCCL world-size-2 tests, the default bench line at N=2 (value, e2e, train + cuda_graph, infer on rank 0)
mkdir -p gpurun_out/n2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_rank_nccl or nccl" > gpurun_out/n2/pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/n2/pytest_nccl.log
tail -3 gpurun_out/n2/pytest_nccl.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2/bench_n2.json 2> gpurun_out/n2/bench_n2.err; echo "bench N=2 rc=$?"
tail -c 400 gpurun_out/n2/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/n2/bench_n2.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
t=d['train']; print('train', t.get('value'), t.get('ms_per_step'), t.get('allreduce',{}).get('exposed_ms'), t.get('cuda_graph'))
print('infer', json.dumps(d['infer'])[:600])
PY
