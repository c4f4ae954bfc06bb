#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 4 gpurun_out/$name.log | cut -c1-600; }
run conv_bo0 300 env AVVAD_SLAB_BO=0 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120 -k "conv and 17-64-64"
run conv_bo1 300 env AVVAD_SLAB_BO=1 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120 -k "conv and 17-64-64"
