#!/bin/bash
# Attention iteration pass on the GPU box: parity tests of the attention kernels, isolated timing, optional timeline.
# Usage (here): gpurun --timeout 900 -- 'bash scripts/gpu_attn_iter.sh <tag> [trace] [probe]'
TAG=${1:-a}
OUT=gpurun_out
mkdir -p $OUT
if [ "$3" = "probe" ]; then ./scripts/ubench/tmem_layout; fi
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention > $OUT/r2_t_$TAG.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/r2_t_$TAG.log
timeout 200 python scripts/gpu_attn_time.py > $OUT/r2_time_$TAG.log 2>&1; cat $OUT/r2_time_$TAG.log
if [ "$2" = "trace" ]; then
  MCA_LIB=$PWD/mca_paper_b200/csrc/libmca_b200_trace.so timeout 300 python scripts/gpu_attn_trace.py > $OUT/r2_trace_$TAG.log 2>&1
  echo "trace rc=$?"; grep -E "^t= ?(5|6|11|12|13) |fwd\]|bwd\]" $OUT/r2_trace_$TAG.log | head -60
fi
