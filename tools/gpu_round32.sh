#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_slab_kernel|tc_tma_kernel" -s 1 -c 5 -o gpurun_out/prof_convs -f python bench.py --ncu --warmup 0 --batch 64 > gpurun_out/ncu_convs.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_convs.log
