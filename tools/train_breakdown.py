#!/usr/bin/env python
"""Where the AV+MCB training step spends its time: CUDA-event brackets around every libavvad call of one step
(avvad.engine functions are wrapped; the step itself is unchanged).  usage: train_breakdown.py [batch ...]"""
import json
import os
import sys
from collections import OrderedDict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "audio-visual-vad_b200")):
    sys.path.insert(0, p)
import torch

from avvad import engine as E, synth
from avvad.train import Trainer

RECS = []


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        RECS.append((label, e0, e1))
        return r
    setattr(obj, name, timed)


def main():
    from packages.models.AV_Net import DeepVAD_AV
    from packages.models import _engine as PE
    dev = torch.device("cuda", 0)
    wrap(E.ResNet18Trunk, "forward_train", "trunk forward (train-mode BN)")
    wrap(E, "mcb_forward_train", "mcb forward")
    wrap(E, "lstm_train_forward", "lstm forward (tape)")
    wrap(E, "batch_bce", "bce loss + dlogits")
    wrap(E, "lstm_train_backward", "lstm backward (BPTT + weight grads)")
    wrap(E, "mcb_backward_bn", "mcb_bn backward")
    wrap(E, "adam_step", "adam")
    wrap(E.Lstm, "load", "lstm weight repack")
    wrap(E.ResNet18Trunk, "load_train", "trunk weight repack")
    out = {}
    for B in [int(a) for a in sys.argv[1:]] or [256, 32]:
        T = 317
        sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), seed=1)
        av = DeepVAD_AV(2, 1024, 1, use_mcb=True)
        av.load_state_dict(sd)
        for q in av.features.parameters():
            q.requires_grad = False
        av = av.to(dev)
        tr = Trainer(av, lr=1e-4)
        a = torch.randn(B, T, 513, device=dev)
        v = torch.randn(B, T, 67, 67, device=dev)
        tgt = (torch.rand(B, T, 1, device=dev) > 0.5).float()
        ln = torch.full((B,), T, dtype=torch.int32, device=dev)
        for _ in range(3):
            tr.step((a, v), tgt, ln)
        torch.cuda.synchronize()
        RECS.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 3
        for _ in range(steps):
            tr.step((a, v), tgt, ln)
        e1.record()
        torch.cuda.synchronize()
        agg = OrderedDict()
        for label, b, e in RECS:
            agg[label] = agg.get(label, 0.0) + b.elapsed_time(e) / steps
        total = e0.elapsed_time(e1) / steps
        agg["(other: autograd glue, arena, zeroing)"] = total - sum(agg.values())
        agg["TOTAL ms/step"] = total
        out[f"B={B}"] = {k: round(v, 3) for k, v in agg.items()}
        del tr, av
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
