#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): the data-parallel step against the oracle on N ranks (both exchange forms at N = 2), the
# multi-GPU pytest, and a short N-rank bench line.  Usage: gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_dp.sh <tag> <N> [modes]'
TAG=${1:-dp}; N=${2:-2}; MODES=${3:-"tiny tiny_z full"}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
for mode in $MODES; do
  port=$((port + 1))
  timeout 900 $RUN --master-port $port scripts/gpu_dp_check.py $mode > $OUT/dp_check_${TAG}_n${N}_$mode.log 2>&1
  echo "dp_check $mode p2p rc=$?"; grep -E "^\[rank 0|DP CHECK|Error|error" $OUT/dp_check_${TAG}_n${N}_$mode.log | cut -c1-260 | head -12
done
port=$((port + 1))
MCA_MULTIMEM=0 timeout 900 $RUN --master-port $port scripts/gpu_dp_check.py tiny > $OUT/dp_check_${TAG}_n${N}_tiny_unicast.log 2>&1
echo "dp_check tiny p2p unicast rc=$?"; grep -E "^\[rank 0|DP CHECK" $OUT/dp_check_${TAG}_n${N}_tiny_unicast.log | cut -c1-260 | head -8
port=$((port + 1))
MCA_P2P=0 timeout 900 $RUN --master-port $port scripts/gpu_dp_check.py tiny > $OUT/dp_check_${TAG}_n${N}_tiny_nccl.log 2>&1
echo "dp_check tiny nccl rc=$?"; grep -E "^\[rank 0|DP CHECK" $OUT/dp_check_${TAG}_n${N}_tiny_nccl.log | cut -c1-260 | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/pytest_multi_$TAG.log
port=$((port + 1))
timeout 600 $RUN --master-port $port bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench n$N rc=$?"; cut -c1-600 $OUT/bench_${TAG}_n$N.json
