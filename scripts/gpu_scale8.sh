#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): the data-parallel step against the oracle on 8 ranks, the N = 8 bench lines (NVSwitch multimem
# optimiser exchange on / off, MMA config, EAO, inference replicas).  Usage: gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale8.sh <tag>'
TAG=${1:-s8}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29701 scripts/gpu_dp_check.py tiny > $OUT/dp_check_${TAG}_n${N}_tiny.log 2>&1
echo "dp_check tiny n$N rc=$?"; grep -E "^\[rank 0|DP CHECK|Error" $OUT/dp_check_${TAG}_n${N}_tiny.log | cut -c1-300 | head -10
timeout 200 $RUN --master-port 29702 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench n$N multimem rc=$?"; grep '^{' $OUT/bench_${TAG}_n$N.json | cut -c1-200
MCA_MULTIMEM=0 timeout 200 $RUN --master-port 29703 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_${TAG}_n${N}_unicast.json 2> $OUT/bench_${TAG}_n${N}_unicast.err
echo "bench n$N unicast rc=$?"; grep '^{' $OUT/bench_${TAG}_n${N}_unicast.json | cut -c1-200
timeout 200 $RUN --master-port 29704 bench.py --gpus $N --steps 30 --warmup 5 --config CMU_config1_z > $OUT/bench_${TAG}_n${N}_mma.json 2> $OUT/bench_${TAG}_n${N}_mma.err
echo "bench n$N MMA rc=$?"; grep '^{' $OUT/bench_${TAG}_n${N}_mma.json | cut -c1-200
timeout 300 $RUN --master-port 29705 bench.py --gpus $N --steps 10 --warmup 3 --config CMU_config1_EAO > $OUT/bench_${TAG}_n${N}_eao.json 2> $OUT/bench_${TAG}_n${N}_eao.err
echo "bench n$N EAO rc=$?"; grep '^{' $OUT/bench_${TAG}_n${N}_eao.json | cut -c1-200; tail -2 $OUT/bench_${TAG}_n${N}_eao.err | cut -c1-200
timeout 200 $RUN --master-port 29706 bench.py --gpus $N --steps 30 --warmup 5 --mode infer > $OUT/bench_${TAG}_n${N}_infer.json 2> $OUT/bench_${TAG}_n${N}_infer.err
echo "bench n$N infer rc=$?"; grep '^{' $OUT/bench_${TAG}_n${N}_infer.json | cut -c1-300
