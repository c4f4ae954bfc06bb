#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | cut -c1-1500; }
run all 900 python -m pytest tests -q -m gpu --timeout 300
run bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
