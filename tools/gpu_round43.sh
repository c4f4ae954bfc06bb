#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_dp_check.py > gpurun_out/train_plain.log 2>&1; echo "exit=$?"; tail -1 gpurun_out/train_plain.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_train.csv python tools/train_dp_check.py > gpurun_out/ncu_train.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/launches_train.csv
