#!/bin/bash
# overlapped host pipeline: parity + bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 8 gpurun_out/$name.log | cut -c1-2500; }
run pipe 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -x
run bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
