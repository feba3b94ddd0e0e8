# 2-GPU validation: NCCL world-size-2 step test, the default bench line at N=2, DDP + CUDA-graph training step
mkdir -p gpurun_out/n2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_rank_nccl" > gpurun_out/n2/pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/n2/pytest_nccl.log
tail -3 gpurun_out/n2/pytest_nccl.log
for mode in "" "--cuda-graph"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --fused-head $mode --steps 20 --warmup 5 > "gpurun_out/n2/train_n2$mode.json" 2> "gpurun_out/n2/train_n2$mode.err"; echo "train $mode rc=$?"
done
grep -iE "Error|error" gpurun_out/n2/train_n2--cuda-graph.err | grep -v Warning | head -5
cat gpurun_out/n2/train_n2*.json | cut -c1-400
