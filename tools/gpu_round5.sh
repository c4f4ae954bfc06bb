#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run plain 600 python bench.py --ncu --warmup 0 --batch 32
run ncu_launches 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_b32_full.csv python bench.py --ncu --warmup 0 --batch 32
