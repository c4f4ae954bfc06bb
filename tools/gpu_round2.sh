#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 12 gpurun_out/$name.log; }
run frontend 300 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_pipeline.py -q -m gpu --timeout 200
run bench 900 python bench.py --steps 3 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
