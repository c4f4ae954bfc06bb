#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_s2d" -s 2 -c 1 -o gpurun_out/prof_stem_s2d -f python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/ncu_stem.log 2>&1
echo "exit=$?"; tail -5 gpurun_out/ncu_stem.log
