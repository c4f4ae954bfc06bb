#!/usr/bin/env python
"""ResNet-18 trunk A/B: device time of one trunk pass and bit-comparison of the features between configurations
(e.g. AVVAD_BLOCK17=0 vs 1).  usage: [env] python tools/micro/trunk_ab.py [frames] [--save f.pt] [--cmp f.pt]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "audio-visual-vad_b200")):
    sys.path.insert(0, p)
import torch

from avvad import engine as E
from avvad import synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 20288
val = lambda k: sys.argv[sys.argv.index(k) + 1] if k in sys.argv else None
sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), 1, "strong")
trunk = E.ResNet18Trunk()
trunk.load(sd, "cuda")
g = torch.Generator().manual_seed(0)
frames = torch.randn(n, 67, 67, generator=g).cuda()
for _ in range(2):
    feat = trunk.forward(frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    feat = trunk.forward(frames)
e1.record()
torch.cuda.synchronize()
tag = " ".join(f"{k[6:]}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("AVVAD_")) or "default"
print(f"[{tag}] trunk pass of {n} frames: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
f = feat[0] if isinstance(feat, (tuple, list)) else feat
f = f.float().cpu()
if val("--save"):
    torch.save(f, val("--save"))
if val("--cmp"):
    r = torch.load(val("--cmp"))
    print(f"[{tag}] vs {val('--cmp')}: rel_fro {((f - r).norm() / r.norm()).item():.3e}, max abs {(f - r).abs().max().item():.3e}, "
          f"bit-identical {bool((f == r).all())}, finite {bool(torch.isfinite(f).all())}", flush=True)
