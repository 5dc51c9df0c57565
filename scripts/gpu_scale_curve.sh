#!/bin/bash
# Same-box scaling curve: bench.py at N = 1, 2, 4 (... up to the GPUs of the box).  Usage: gpurun --gpus G -- 'bash scripts/gpu_scale_curve.sh <tag> <G>'
TAG=${1:-sc}; G=${2:-4}
OUT=gpurun_out; mkdir -p $OUT
n=1; port=29960
while [ $n -le $G ]; do
  port=$((port + 1))
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 30 --warmup 5 > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
  fi
  echo "n=$n rc=$? $(grep -o '"value": [0-9.]*' $OUT/bench_${TAG}_n$n.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' $OUT/bench_${TAG}_n$n.json | head -1)"
  n=$((n * 2))
done
