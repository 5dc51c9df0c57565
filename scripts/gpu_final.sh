#!/bin/bash
# Round-end GPU pass: parity tests, smoke, the bench line (N=1), the reference arm, the ncu launch list of exactly ONE step
# (bench.py --profile-step brackets it with cudaProfilerStart/Stop) and full ncu captures of the attention kernels, the GEMM
# family and the bandwidth kernels of that step.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_final.sh <tag>'
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
MCA_BENCH_TABLE=$OUT/kernel_table_$TAG.json python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_ref_$TAG.json
for spec in "tcga:--config TCGA_config1 --variant tcga" "d40:--config CMU_config1_d40 --variant dropout_ragged" \
            "d40fast:--config CMU_config1_d40 --variant dropout_ragged --varlen fast" "mma:--config CMU_config1_z" \
            "eao:--config CMU_config1_EAO --steps 10"; do
  name=${spec%%:*}; flags=${spec#*:}
  MCA_BENCH_TABLE=$OUT/kernel_table_${TAG}_$name.json python bench.py --steps 20 --warmup 5 --no-cpu-baseline $flags > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  echo "bench $name rc=$?"; cut -c1-200 $OUT/bench_${TAG}_$name.json
done
python bench.py --mode infer --steps 30 --warmup 5 > $OUT/bench_${TAG}_infer.json 2> $OUT/bench_${TAG}_infer.err; echo "infer rc=$?"; cut -c1-200 $OUT/bench_${TAG}_infer.json
STEP="python bench.py --profile-step --no-graphs --warmup 3 --no-cpu-baseline"
$STEP > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_$TAG.csv \
    $STEP > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
for spec in "attn:attn_fwd_kernel|attn_bwd_kernel:10" "gemm:gemm2_tc_kernel|gemm_tc_kernel:60" "bw:ln512|add_ln512|clip_adamw|pool_fwd_cl|pool_bwd_cl|loss_:30"; do
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$regex" -c $cnt -f -o $OUT/prof_${TAG}_$name \
      $STEP > $OUT/ncu_full_${TAG}_$name.log 2>&1
  echo "ncu full $name rc=$?"
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page raw --csv > $OUT/prof_${TAG}_${name}_raw.csv 2>/dev/null
  if [ "$name" = "attn" ]; then
    ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/prof_${TAG}_${name}_source.csv.gz
  fi
  rm -f $OUT/prof_${TAG}_$name.ncu-rep
done
du -sh $OUT
