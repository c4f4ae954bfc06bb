#!/bin/bash
# 8-warp prefetching epilogue + K-concatenated downsample branch: parity + bench (fused vs unfused)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 8 gpurun_out/$name.log | cut -c1-400; }
run gemm 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 300 -x
run models 600 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 300
run bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
AVVAD_FUSE_DS=0 run bench_nofuse 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
AVVAD_PROFILE_PER_LAUNCH=1 AVVAD_LAYER_DUMP=gpurun_out/layers.json run bench_layers 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run all 900 python -m pytest tests -q -m gpu --timeout 300
