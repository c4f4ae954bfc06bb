#!/usr/bin/env python
"""Multi-GPU check of the training path (run under torchrun, one rank per GPU):
  1. audio-only model: all-reduced gradients of a batch sharded over the ranks == single-GPU gradients of the whole
     batch (the loss is a sum over utterances, no BatchNorm -> exact up to fp summation order);
  2. AV (MCB) training step timing, batch 256 split over the ranks (BASELINE config 5), device-timed, max over ranks.
Prints one JSON line on rank 0."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "audio-visual-vad_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from avvad import engine as E, synth
from avvad.train import Trainer, allreduce_gradients


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from packages.models.Audio_Net import DeepVAD_audio
    from packages.models.AV_Net import DeepVAD_AV

    out = {"world": world}
    quick = os.environ.get("AVVAD_TRAIN_QUICK") == "1"  # profiling: one warm-up + one timed AV step only
    # ---- 1. gradient equivalence (audio-only)
    B, T = 8 * world, 40
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, 513, generator=g)
    y = (torch.rand(B, T, 1, generator=g) > 0.5).float()
    lens = [T - (i * 3) % 17 for i in range(B)]
    sd = synth.seeded_state_dict(synth.model_spec("audio"), seed=5)
    m = DeepVAD_audio(2, 1024, 1)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    sl = slice(rank * B // world, (rank + 1) * B // world)
    logits = m(x[sl].to(dev), lens[sl])
    _, _, dl = E.batch_bce(logits, y[sl].to(dev), lens[sl], 1e-8, want_grad=True)
    logits.backward(dl)
    allreduce_gradients(list(m.parameters()))
    sharded = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()
    logits = m(x.to(dev), lens)
    _, _, dl = E.batch_bce(logits, y.to(dev), lens, 1e-8, want_grad=True)
    logits.backward(dl)
    worst = 0.0
    for k, p in m.named_parameters():
        e = ((sharded[k] - p.grad).norm() / (p.grad.norm() + 1e-30)).item()
        worst = max(worst, e)
    out["sharded_vs_full_grad_rel_err"] = worst

    # ---- 2. AV training step timing (global batch 256)
    Bg, Tt = 256, 317
    Bl = Bg // world
    sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=True), seed=1)
    av = DeepVAD_AV(2, 1024, 1, use_mcb=True)
    av.load_state_dict(sd)
    for q in av.features.parameters():
        q.requires_grad = False
    av = av.to(dev)
    tr = Trainer(av, lr=1e-4)
    a = torch.randn(Bl, Tt, 513, device=dev)
    v = torch.randn(Bl, Tt, 67, 67, device=dev)
    tgt = (torch.rand(Bl, Tt, 1, device=dev) > 0.5).float()
    ln = torch.full((Bl,), Tt, dtype=torch.int32, device=dev)
    for _ in range(1 if quick else 2):
        loss = tr.step((a, v), tgt, ln)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 1 if quick else 3
    for _ in range(steps):
        loss = tr.step((a, v), tgt, ln)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out["av_train_step_ms"] = float(ms)
    out["av_train_frames_per_s"] = Bg * Tt / (float(ms) / 1e3)
    out["loss"] = float(loss)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
