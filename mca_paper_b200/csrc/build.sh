#!/bin/bash
# Builds libmca_b200.so for sm_100a (nvcc cross-compiles without a GPU).  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../include -Xcompiler -fPIC --threads 0"
mkdir -p build
pids=()
for f in *.cu; do
  o="build/${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ ptx.cuh -nt "$o" ] || [ gemm_epilogue.cuh -nt "$o" ] || [ runtime.h -nt "$o" ] || [ ../../include/mca_b200.h -nt "$o" ]; then
    nvcc $FLAGS ${MCA_NVCC_EXTRA} -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmca_b200.so build/*.o
echo "built $(pwd)/libmca_b200.so"
