#!/bin/bash
# Debug build with the clock64 timeline of the attention backward (-DMCA_TRACE) -> csrc/libmca_b200_trace.so
# (used as MCA_LIB=mca_paper_b200/csrc/libmca_b200_trace.so python scripts/gpu_attn_trace.py)
set -e
cd "$(dirname "$0")/../mca_paper_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../include -Xcompiler -fPIC"
mkdir -p build_trace
nvcc $FLAGS -DMCA_TRACE -c attention_bwd.cu -o build_trace/attention_bwd.o
objs=""
for f in build/*.o; do
  [ "$(basename $f)" = "attention_bwd.o" ] || objs="$objs $f"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmca_b200_trace.so build_trace/attention_bwd.o $objs
echo "built libmca_b200_trace.so"
