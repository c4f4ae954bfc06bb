#!/bin/bash
# Multi-GPU: inference bench (utterance-sharded, no collective) and the data-parallel training step, N ranks.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu/scaling.sh 2'
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g$N.log 2>&1; echo "bench exit=$?"
grep '^{' gpurun_out/bench_g$N.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    tools/train_dp_check.py > gpurun_out/train_g$N.log 2>&1; echo "train exit=$?"
grep '^{' gpurun_out/train_g$N.log
