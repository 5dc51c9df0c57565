#!/bin/bash
# Round-end GPU pass: parity tests, the bench line (N=1), the reference arm, the ncu launch list of one step and full
# ncu captures of the attention kernels and of the GEMM family.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_final.sh <tag>'
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
MCA_BENCH_TABLE=$OUT/kernel_table_$TAG.json python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_ref_$TAG.json
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline"
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 420 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH_SHORT > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
for spec in "attn:attn_fwd_kernel|attn_bwd_kernel:10:10" "gemm:gemm2_tc_kernel|gemm_tc_kernel:80:40" "ln:ln512|add_ln512|clip_adamw:30:12"; do
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; rest=${rest#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c $cnt -f -o $OUT/prof_${TAG}_$name \
      $BENCH_SHORT > $OUT/ncu_full_${TAG}_$name.log 2>&1
  echo "ncu full $name rc=$?"
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page raw --csv > $OUT/prof_${TAG}_${name}_raw.csv 2>/dev/null
  if [ "$name" = "attn" ]; then
    ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/prof_${TAG}_${name}_source.csv.gz
  fi
  rm -f $OUT/prof_${TAG}_$name.ncu-rep
done
du -sh $OUT
