# timing experiments on K4w / K4x (library built with -DIHPR_TIMING_EXPERIMENTS; results are WRONG by design when a knob is set)
mkdir -p gpurun_out/r2d
for dbg in 0 1 2 3 4 5 8 15; do
  echo "== IHPR_K4_DEBUG=$dbg" >> gpurun_out/r2d/k4_knockout.txt
  IHPR_K4_DEBUG=$dbg timeout 200 python tools/head_bench.py --B 32 --iters 10 2>&1 | head -1 >> gpurun_out/r2d/k4_knockout.txt
done
cat gpurun_out/r2d/k4_knockout.txt
