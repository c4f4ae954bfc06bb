#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 4 gpurun_out/$name.log | cut -c1-300; }
run models 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_gemm.py -q -m gpu --timeout 300
AVVAD_LAYER_DUMP=gpurun_out/layers.json run bench 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
