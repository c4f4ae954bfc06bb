#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log; }
run frontend 300 python -m pytest tests/test_gpu_frontend.py -q -m gpu --timeout 200
AVVAD_LAYER_DUMP=gpurun_out/layers.json run bench 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
cat gpurun_out/layers.json
run bench_ncu_plain 600 python bench.py --ncu --warmup 1 --batch 32
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_b32.csv python bench.py --ncu --warmup 1 --batch 32
