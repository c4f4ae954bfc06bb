#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/train_dp_check.py > gpurun_out/train_g2.log 2>&1; echo "g2 exit=$?"; grep '^{' gpurun_out/train_g2.log
timeout 600 python tools/train_dp_check.py > gpurun_out/train_g1.log 2>&1; echo "g1 exit=$?"; grep '^{' gpurun_out/train_g1.log
