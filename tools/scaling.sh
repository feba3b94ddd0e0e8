#!/bin/bash
# tools/scaling.sh -- 1/2/4/8-GPU runs of the two bench workloads on one box (BASELINE.json configs 2-4); writes gpurun_out/scaling.jsonl
mkdir -p gpurun_out
out=gpurun_out/scaling.jsonl; : > $out
run() { # N, args...
  n=$1; shift
  if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 "$@" 2>/dev/null | tail -1 >> $out
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" 2>/dev/null | tail -1 >> $out; fi
}
for n in 1 2 4 8; do run $n --workload train --fused-head --steps 20 --warmup 5; done
for n in 1 8; do run $n --workload train --steps 20 --warmup 5; done
run 8 --workload train --fused-head --resnet 152 --batch 64 --steps 10 --warmup 3
run 1 --workload train --fused-head --resnet 152 --batch 64 --steps 10 --warmup 3
for n in 2 4 8; do run $n --steps 30 --warmup 5 --no-cpu --no-e2e; done
python - <<'PY'
import json
for l in open("gpurun_out/scaling.jsonl"):
    try: d = json.loads(l)
    except Exception: print("BAD", l[:200]); continue
    print(d["metric"], "| N =", d["n_gpus"], "|", d["config"].get("workload", "")[:60], "|", d["config"].get("criterion", ""), "| value %.0f %s | %.3f ms/step" % (d["value"], d["unit"], d["ms_per_step"]))
PY
