#!/bin/bash
# Builds the phase-timer variant of the library (-DTDB_TC_TIMING) into torch_de_solver_b200/lib_timing/; use it with
#   TDB200_LIB=$PWD/torch_de_solver_b200/lib_timing/libtedeous_b200.so TDB200_TC_TIMING=1 python bench.py ...
set -e
cd "$(dirname "$0")/.."
OUT=torch_de_solver_b200/lib_timing
mkdir -p $OUT
pids=()
for f in torch_de_solver_b200/csrc/*.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DTDB_TC_TIMING -I include \
       -I torch_de_solver_b200/csrc -c $f -o $OUT/$(basename ${f%.cu}).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT/libtedeous_b200.so $OUT/*.o -cudart static
echo built $OUT/libtedeous_b200.so
