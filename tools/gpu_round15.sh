#!/bin/bash
mkdir -p gpurun_out
for sub in 128 256 512; do
  AVVAD_STEM_SUB=$sub timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sub$sub.log 2>&1
  echo "sub=$sub exit=$?"
done
