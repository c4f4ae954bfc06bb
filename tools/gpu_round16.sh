#!/bin/bash
mkdir -p gpurun_out
python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_b32_v3.csv python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/ncu_l.log 2>&1
echo done
