#!/bin/bash
mkdir -p gpurun_out
AVVAD_SLAB=2 AVVAD_LAYER_DUMP=gpurun_out/layers_slab2.json timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_slab2.log 2>&1
AVVAD_LAYER_DUMP=gpurun_out/layers.json timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo done
