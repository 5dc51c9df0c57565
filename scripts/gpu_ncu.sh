#!/bin/bash
# Full ncu capture of selected kernels of the bench step, exported to CSV on the box (gpurun_out/ is capped at 64 MiB).
# usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex> <skip> <count>
TAG=$1; KREGEX=$2; SKIP=${3:-0}; COUNT=${4:-4}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s $SKIP -c $COUNT -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/prof_${TAG}_details.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/prof_${TAG}_source.csv 2>/dev/null
gzip -f $OUT/prof_${TAG}_source.csv
sz=$(stat -c %s $OUT/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f $OUT/prof_$TAG.ncu-rep; fi
du -sh $OUT
