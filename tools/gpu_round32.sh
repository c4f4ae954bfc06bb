#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_slab_kernel" -s 1 -c 1 -o gpurun_out/prof_slab2 -f python bench.py --ncu --warmup 0 --batch 64 > gpurun_out/ncu_slab2.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_slab2.log
