#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 35 -c 19 -o gpurun_out/prof_trunk python bench.py --ncu --warmup 0 --batch 32 > gpurun_out/ncu_full.log 2>&1
echo "exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
