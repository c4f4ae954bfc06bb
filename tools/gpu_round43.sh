#!/bin/bash
mkdir -p gpurun_out
AVVAD_TRAIN_QUICK=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_train.csv python tools/train_dp_check.py > gpurun_out/ncu_train.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/launches_train.csv
