#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 4 gpurun_out/$name.log | cut -c1-300; }
run models 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_pipeline.py tests/test_gpu_gemm.py -q -m gpu --timeout 300
AVVAD_LAYER_DUMP=gpurun_out/layers.json run bench 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
AVVAD_SLAB=2 run conv_all 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py -q -m gpu --timeout 120
AVVAD_SLAB=2 AVVAD_LAYER_DUMP=gpurun_out/layers_slab2.json run bench_slab2 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
