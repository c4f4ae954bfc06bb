#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --ncu --warmup 0 > gpurun_out/ncu_launches.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_r01b.csv
