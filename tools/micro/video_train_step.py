#!/usr/bin/env python
"""scripts/train_video_net.py's step at its own batch size (16 utterances, ~300 frames each = 4,800 frames), trunk
TRAINABLE: torch.optim.Adam over all 25.9 M parameters, the script's loss loop, loss.backward().  Device time per step and
the split between forward and backward."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "audio-visual-vad_b200")):
    sys.path.insert(0, p)
import torch

from avvad import synth
from packages.models.Video_Net import DeepVAD_video
from packages.models.utils import binary_cross_entropy

B, T = int(os.environ.get("B", 16)), 300
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, 67, 67, generator=g).cuda()
y = (torch.rand(B, T, 1, generator=g) > 0.5).long().cuda()
lens = torch.tensor([T - (7 * i) % 40 for i in range(B)])
m = synth.fill_module_(DeepVAD_video(2, 1024, 1), seed=3).cuda()
opt = torch.optim.Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.999))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]


def step(timed=False):
    m.train()
    if timed: ev[0].record()
    out = m(x, lens.cuda())
    loss = 0.
    for length, pred, target in zip(lens, out, y):
        loss += binary_cross_entropy(pred[:length], target[:length], 1e-8)
    if timed: ev[1].record()
    loss.backward()
    if timed: ev[2].record()
    opt.step()
    opt.zero_grad()
    if timed: ev[3].record()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
fw = bw = op = 0.0
N = 5
for _ in range(N):
    loss = step(True)
    torch.cuda.synchronize()
    fw += ev[0].elapsed_time(ev[1]); bw += ev[1].elapsed_time(ev[2]); op += ev[2].elapsed_time(ev[3])
print(f"video-net training step, B={B} x T={T} ({B * T} frames, trunk trainable, {sum(p.numel() for p in m.parameters()) / 1e6:.1f} M "
      f"parameters): forward + loss {fw / N:.1f} ms, backward {bw / N:.1f} ms, torch Adam {op / N:.1f} ms = "
      f"{(fw + bw + op) / N:.1f} ms/step = {B * T / ((fw + bw + op) / N) * 1e3 / 1e3:.1f} k frames/s; loss {loss.item():.3f}; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
