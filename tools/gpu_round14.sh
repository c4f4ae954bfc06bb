#!/bin/bash
mkdir -p gpurun_out
python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_tma_kernel -s 3 -c 1 -o gpurun_out/prof_stemgemm python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/ncu_stem.log 2>&1
tail -2 gpurun_out/ncu_stem.log
