#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_models.py -q -m gpu --timeout 200 -x 2>&1 | tail -8
timeout 600 python tools/train_dp_check.py > gpurun_out/train_plain.log 2>&1; echo "exit=$?"; tail -1 gpurun_out/train_plain.log
