#!/bin/bash
mkdir -p gpurun_out
python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_tma_kernel|tc_slab_kernel|stem_fused_kernel|lstm_persist_kernel" -s 21 -c 22 -o gpurun_out/prof_r01_final python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
