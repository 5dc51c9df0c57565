#!/bin/bash
# One GPU-box pass: parity tests, the bench line, the per-launch ncu list and full ncu captures of the top kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag> [full-capture regex]'
TAG=${1:-r1}
KREGEX=${2:-"attn_fwd_kernel|attn_bwd|gemm_tc_kernel"}
OUT=gpurun_out
mkdir -p $OUT
set -o pipefail
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
MCA_BENCH_TABLE=$OUT/kernel_table_$TAG.json python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline"
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 420 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH_SHORT > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
if [ -n "$KREGEX" ]; then
  $BENCH_SHORT > $OUT/plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s ${NCU_SKIP:-130} -c ${NCU_COUNT:-40} -f -o $OUT/prof_$TAG \
      $BENCH_SHORT > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full rc=$?"
  # gpurun_out/ is capped at 64 MiB: export what is read offline as CSV, keep the .ncu-rep only when small
  ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/prof_${TAG}_details.csv 2>/dev/null
  ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/prof_${TAG}_source.csv 2>/dev/null
  gzip -f $OUT/prof_${TAG}_source.csv
  sz=$(stat -c %s $OUT/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 30000000 ]; then rm -f $OUT/prof_$TAG.ncu-rep; echo "dropped .ncu-rep ($sz bytes)"; fi
fi
du -sh $OUT
ls -la $OUT
