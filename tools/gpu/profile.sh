#!/bin/bash
# Profiles for profiles/: (1) ncu launch list of the bench command, (2) ncu --set full of one 64-utterance pass.
# Each capture runs only after the same command exited 0 without ncu.
#   gpurun --timeout 2400 -- 'bash tools/gpu/profile.sh'
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-config4 > gpurun_out/bench_plain.log 2>&1 || exit 1
AVVAD_LSTM_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --ncu --warmup 0 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
AVVAD_LSTM_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"tc_slab_kernel|tc_block17|tc_tma_kernel|stem_s2d|lstm_persist|lstm_pair|mcb_row|mcb_apply|frontend_kernel|frontend_reg" -c 32 \
    -o gpurun_out/prof_full -f python bench.py --ncu --warmup 0 --batch 64 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
# the CTA-pair LSTM recurrence at the benchmark batch (one launch per layer: AVVAD_LSTM_CHUNKS=1)
AVVAD_LSTM_COOP=0 AVVAD_LSTM_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"lstm_pair" -c 2 -o gpurun_out/prof_lstm_pair -f python bench.py --ncu --warmup 0 > gpurun_out/ncu_lstm_pair.log 2>&1
echo "lstm pair capture exit=$?"
# gpurun copies at most 64 MiB back: export the raw metric pages here and drop the large report
ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_lstm_pair.ncu-rep --page raw --csv > gpurun_out/prof_lstm_pair_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
[ "$(stat -c %s gpurun_out/prof_full.ncu-rep)" -gt 40000000 ] && rm -f gpurun_out/prof_full.ncu-rep
du -sh gpurun_out
