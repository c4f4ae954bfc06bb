#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_persist" -c 1 -o gpurun_out/prof_lstm -f python bench.py --ncu --warmup 0 --batch 256 > gpurun_out/ncu_lstm.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_lstm.log
