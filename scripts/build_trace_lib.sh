#!/bin/bash
# Debug build with the clock64 timelines of the attention kernels (-DMCA_TRACE) -> csrc/libmca_b200_trace.so
# (used as MCA_LIB=mca_paper_b200/csrc/libmca_b200_trace.so python scripts/gpu_attn_trace.py)
set -e
cd "$(dirname "$0")/../mca_paper_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../include -Xcompiler -fPIC"
mkdir -p build_trace
objs=""
for f in attention_bwd attention_fwd; do
  nvcc $FLAGS -DMCA_TRACE -c $f.cu -o build_trace/$f.o &
done
wait
for f in build/*.o; do
  case "$(basename $f)" in attention_bwd.o|attention_fwd.o) ;; *) objs="$objs $f";; esac
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmca_b200_trace.so build_trace/attention_bwd.o build_trace/attention_fwd.o $objs
echo "built libmca_b200_trace.so"
