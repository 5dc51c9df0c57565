#!/bin/bash
# Short GPU pass: parity tests, a 10-step bench with the per-kernel table, optional attention-backward timeline.
# Usage (here): gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag> [trace]'
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_$TAG.log
MCA_BENCH_TABLE=$OUT/kernel_table_$TAG.json python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
python scripts/show_table.py $OUT/kernel_table_$TAG.json | head -40
if [ "$2" = "trace" ]; then
  MCA_LIB=$PWD/mca_paper_b200/csrc/libmca_b200_trace.so python scripts/gpu_attn_trace.py > $OUT/trace_$TAG.log 2>&1
  echo "trace rc=$?"; head -45 $OUT/trace_$TAG.log
fi
