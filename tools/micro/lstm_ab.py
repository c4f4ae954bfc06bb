#!/usr/bin/env python
"""LSTM recurrence A/B: device time of avvad_lstm_forward (2 x LSTM-1024 + head) at the benchmark shape, the output of the
current configuration against the one-CTA-per-block kernel (AVVAD_LSTM_PAIR=0 in a child process), and optionally a
per-step timeline of the CTA-pair kernel.
usage: [AVVAD_LSTM_PAIR=0|1] python tools/micro/lstm_ab.py [B] [T] [--trace] [--train] [--save f.pt] [--cmp f.pt]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "audio-visual-vad_b200")):
    sys.path.insert(0, p)
import torch

from avvad import engine as E
from avvad import lib as L
from avvad import synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 256
T = int(args[1]) if len(args) > 1 else 317
opt = lambda k: k in sys.argv
val = lambda k: sys.argv[sys.argv.index(k) + 1] if k in sys.argv else None

sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), 1, "strong")
lstm = E.Lstm(2, 1024, 1024, 1)
lstm.load(sd, "cuda", "lstm_merged", "vad_merged")
g = torch.Generator().manual_seed(0)
x = lstm.new_input(B, T, "cuda")
x[:, :, :1024] = (torch.randn(B, T, 1024, generator=g) * 0.5).to(torch.bfloat16).cuda()
lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g).tolist()
lens[0] = T


def run():
    if opt("--train"):
        lg, tape = E.lstm_train_forward(lstm, x, lens)
        lstm._tape_busy = False
        return lg
    return lstm.forward(x, lens)[0]


for _ in range(3):
    out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
E.profile_enable(True)
E.profile_clear()
run()
rec_ms, _, rec_n = E.profile_read(2)
E.profile_enable(False)
tag = " ".join(f"{k[11:]}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("AVVAD_LSTM_")) or "default"
print(f"[{tag}] B={B} T={T} train={opt('--train')}: forward {ms:.3f} ms; recurrence kernels {rec_ms:.3f} ms in {rec_n} "
      f"launches = {rec_ms * 1e3 / (2 * T):.2f} us per time step and layer", flush=True)
valid = torch.zeros(B, T, dtype=torch.bool)
for b, l in enumerate(lens):
    valid[b, :l] = True
if val("--save"):
    torch.save(out.cpu(), val("--save"))
if val("--cmp"):
    ref = torch.load(val("--cmp"))
    a, r = out.cpu()[valid], ref[valid]
    print(f"[{tag}] vs {val('--cmp')}: logits rel_fro {((a - r).norm() / r.norm()).item():.3e}, max abs "
          f"{(a - r).abs().max().item():.3e}, bit-identical {bool((out.cpu() == ref).all())}", flush=True)

if opt("--trace"):
    n_cta, NS = 128, 12
    buf = torch.zeros(n_cta * T * NS, dtype=torch.int64, device="cuda")
    L.lib().avvad_debug_lstm_trace(L.ptr(buf))
    run()
    torch.cuda.synchronize()
    L.lib().avvad_debug_lstm_trace(None)
    tr = buf.cpu().view(n_cta, T, NS).double()
    names = ["kb0 flags seen", "kb0 TMA issued", "last TMA issued", "first MMA", "last commit", "acc seen", "tmem read",
             "h stored", "proxy fence", "CTA barrier", "flag out"]
    # the second layer overwrote the first: one layer's timeline.  Steps 20..T-20, relative to the CTA's previous flag
    for cta in (0, 1, 62, 63):
        t0 = tr[cta, 20:T - 20]
        prev_flag = tr[cta, 19:T - 21, 10]
        line = []
        for s_ in range(11):
            if (cta % 2 == 1 and s_ in (3, 4)) or s_ == 8:
                continue
            d = (t0[:, s_] - prev_flag)
            line.append(f"{names[s_]} {d.mean().item() / 1e3:5.2f}")
        per = (tr[cta, 21:T - 19, 10] - tr[cta, 20:T - 20, 10]).mean().item() / 1e3
        print(f"CTA {cta:2d}: period {per:5.2f} us | us since own previous flag: " + " | ".join(line), flush=True)
    tr = tr[: (128 if os.environ.get("AVVAD_LSTM_NP") == "64" else 64)]
    fl = tr[:, 20:T - 20, 10]
    print(f"flag-out spread over the CTAs per step: mean {(fl.max(0).values - fl.min(0).values).mean().item() / 1e3:.2f} us",
          flush=True)
    # visibility: flags of pairs 0,1 (same rank) at step t -> "kb0 flags seen" of step t+1
    for r in (0, 1):
        npk = 4 if os.environ.get("AVVAD_LSTM_NP") == "64" else 2
        src = torch.stack([tr[2 * p_ + r, 20:T - 20, 10] for p_ in range(npk)]).max(0).values
        seen = tr[r::2, 21:T - 19, 0]
        d = seen - src[None, :]
        print(f"rank {r}: flags of pairs 0,1 out -> seen by the consumers' pollers: mean {d.mean().item() / 1e3:.2f} us, "
              f"min {d.min().item() / 1e3:.2f}, max over CTAs (mean over steps) {d.max(0).values.mean().item() / 1e3:.2f}",
              flush=True)
