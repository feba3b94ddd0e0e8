#!/bin/bash
# tools/scaling8.sh -- final-kernel refresh of the 8-GPU lines (train fused head, hot path); writes gpurun_out/scaling8.jsonl
mkdir -p gpurun_out; out=gpurun_out/scaling8.jsonl; : > $out
run() { n=$1; shift
  if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 "$@" 2>/dev/null | tail -1 >> $out
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n "$@" 2>/dev/null | tail -1 >> $out; fi; }
run 8 --workload train --fused-head --steps 30 --warmup 5
run 4 --workload train --fused-head --steps 30 --warmup 5
run 8 --workload train --fused-head --resnet 152 --batch 64 --steps 10 --warmup 3
run 8 --steps 50 --warmup 5 --no-cpu --no-e2e
run 4 --steps 50 --warmup 5 --no-cpu --no-e2e
python - <<'PY'
import json
for l in open("gpurun_out/scaling8.jsonl"):
    try: d = json.loads(l)
    except Exception: print("BAD", l[:200]); continue
    print(d["metric"], "| N =", d["n_gpus"], "|", d["config"].get("workload", "")[:70], "|", d["config"].get("criterion", ""), "| value %.0f %s | %.3f ms/step" % (d["value"], d["unit"], d["ms_per_step"]))
PY
