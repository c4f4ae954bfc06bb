#!/bin/bash
# image-as-operand stem: parity, then bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log | cut -c1-1800; }
run stem 300 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 120 -x -k "stem or u8_source"
run models 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300
run bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
