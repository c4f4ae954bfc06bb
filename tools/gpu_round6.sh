#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | cut -c1-200; }
run models 600 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 300 -k trunk
AVVAD_LAYER_DUMP=gpurun_out/layers_cg.json run bench_cg 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
AVVAD_CA=1 AVVAD_LAYER_DUMP=gpurun_out/layers_ca.json run bench_ca 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
