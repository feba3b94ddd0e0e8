mkdir -p gpurun_out/r2g
timeout 300 python -m pytest tests/test_gpu_deconv.py -m gpu -q -x -k graphed 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --train-steps 5 > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err; echo rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2g/bench.json') if l.startswith('{')][-1])
print(json.dumps(d['infer'])[:1500])
PY
