#!/bin/bash
# WaveNet encoder: fused kernel vs per-layer path (device time), then one ncu --set full capture of the fused kernel
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
AVVAD_WAVENET_FUSED=1 timeout 300 python tools/micro/wavenet_ab.py 2>&1 | tail -1
AVVAD_WAVENET_FUSED=0 timeout 300 python tools/micro/wavenet_ab.py 2>&1 | tail -1
AVVAD_WAVENET_FUSED=1 timeout 600 ncu --set full --clock-control none -k regex:"wavenet_fused" -c 1 -o gpurun_out/prof_wavenet -f \
    python tools/micro/wavenet_ab.py > gpurun_out/ncu_wavenet.log 2>&1
echo "capture exit=$?"
ncu -i gpurun_out/prof_wavenet.ncu-rep --page raw --csv > gpurun_out/prof_wavenet_raw.csv 2>/dev/null
