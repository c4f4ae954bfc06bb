#!/bin/bash
# Profiles for profiles/: (1) ncu launch list of the bench command, (2) ncu --set full of one 64-utterance pass.
# Each capture runs only after the same command exited 0 without ncu.
#   gpurun --timeout 2400 -- 'bash tools/gpu/profile.sh'
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --ncu --warmup 0 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"tc_slab_kernel|tc_tma_kernel|stem_s2d|lstm_persist|mcb_row|frontend_kernel" -c 24 \
    -o gpurun_out/prof_full -f python bench.py --ncu --warmup 0 --batch 64 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
