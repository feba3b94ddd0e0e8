#!/bin/bash
# tools/build_experiment_lib.sh -- the library with -DIHPR_TIMING_EXPERIMENTS (knock-out / policy knobs of K4, K5, K9: WRONG results by design when a
# knob is set) into build/lib_exp/, next to the release build.  Run in the container (nvcc cross-compiles); the .so travels to the GPU box with the
# snapshot, and tools/k5_experiments.sh puts it first on LD_LIBRARY_PATH.
set -e
cd "$(dirname "$0")/.."
SRC=integral-human-pose-regression-for-3d-human-pose-estimation_b200/csrc
mkdir -p build/lib_exp build/obj_exp
for f in $SRC/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DIHPR_TIMING_EXPERIMENTS -c "$f" -o "build/obj_exp/$(basename "${f%.cu}").o" &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/lib_exp/libihpr_b200.so build/obj_exp/*.o
ls -la build/lib_exp/libihpr_b200.so
