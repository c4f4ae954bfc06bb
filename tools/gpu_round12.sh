#!/bin/bash
mkdir -p gpurun_out
python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_slab_kernel -s 5 -c 2 -o gpurun_out/prof_slab python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/ncu_slab.log 2>&1
tail -2 gpurun_out/ncu_slab.log
