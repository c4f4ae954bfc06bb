#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -s --timeout 300 > gpurun_out/train.log 2>&1; echo "exit=$?"
tail -30 gpurun_out/train.log | cut -c1-500
