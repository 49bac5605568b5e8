#!/bin/bash
# Multi-GPU evidence: bench.py at N GPUs (headline config), optionally cfg4 (100M rows, no rerank).
#   gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_scale.sh N r1 [cfg4]'
n=$1; tag=${2:-r1}; out=gpurun_out; mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port"
timeout 300 $T 29551 bench.py --gpus $n > $out/bench_${tag}_n$n.json 2> $out/bench_${tag}_n$n.err; echo "bench exit=$?"
cat $out/bench_${tag}_n$n.json
if [ "$3" = "cfg4" ]; then
  timeout 600 $T 29552 bench.py --gpus $n --rows 100000000 --no-rerank --steps 5 > $out/bench_${tag}_cfg4_n$n.json 2> $out/bench_${tag}_cfg4_n$n.err; echo "cfg4 exit=$?"
  cat $out/bench_${tag}_cfg4_n$n.json
fi
