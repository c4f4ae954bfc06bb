#!/usr/bin/env python
"""Fused WaveNet stack (csrc/wavenet_fused.cuh) vs the per-layer path (AVVAD_WAVENET_FUSED=0): device time of one encode.
usage: AVVAD_WAVENET_FUSED=1|0 python tools/micro/wavenet_ab.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "audio-visual-vad_b200")):
    sys.path.insert(0, p)
import torch

from avvad import synth
from packages.models.wavenet_autoencoder import wavenet_autoencoder

B, N = 64, 16000
dil = [1, 2, 4, 8, 1, 2, 4, 8]
wn = wavenet_autoencoder(2, 16, dil, 64, 64, 64, 50, use_bias=True)
synth.fill_module_(wn, seed=1)
wn = wn.cuda().eval()
x = torch.randn(B, 16, N, device="cuda")
with torch.no_grad():
    for _ in range(3):
        wn(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        wn(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
macs = B * (N - 31) * (2 * 16 * 64 + 8 * (2 * 64 * 64 + 64 * 64) + 64 * 64)
print(f"AVVAD_WAVENET_FUSED={os.environ.get('AVVAD_WAVENET_FUSED', '1')}: {ms:.3f} ms per encode of B={B} x N={N} "
      f"(8 dilated layers, 64 channels): {B * N / ms / 1e3:.1f} M samples/s, {2 * macs / ms / 1e9:.1f} TFLOP/s algorithmic")
