"""Data-parallel training step (BASELINE config 5): one process per GPU, local batch B/N, identical replicas,
and ONE all-reduce (sum) of the trainable gradients per step -- what remains of nn.DataParallel's per-step parameter
broadcast / scatter / gather / reduce-add (scripts/train_AV_net.py:193,293-307).

The loss is a SUM over utterances (scripts/train_AV_net.py:298-302), so gradients add across ranks without rescaling.
BatchNorm batch statistics and the MCB whole-tensor L2 norm stay per rank, exactly as they are per DataParallel
replica in the reference.  torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is plumbing only.

Gradients live in ONE flat fp32 arena: every parameter's ``.grad`` is a view into it, so the all-reduce runs in place on
the arena (no gather into a bucket, no scatter back) and zeroing the gradients is one memset.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientArena:
    """One contiguous fp32 buffer holding the gradients of `params` (in order); ``p.grad`` are views into it."""

    def __init__(self, params: Iterable[torch.Tensor]):
        self.params: List[torch.Tensor] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self):
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = self.flat[off:off + k].view_as(p)
            off += k

    def attached(self) -> bool:
        off = 0
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + 4 * off or g.shape != p.shape:
                return False
            off += p.numel()
        return True

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        """Sum over all ranks, in place on the arena (a no-op for a single process)."""
        if not self.attached():   # somebody replaced a .grad (zero_grad(set_to_none=True), p.grad = ...): re-gather
            off = 0
            for p in self.params:
                k = p.numel()
                if p.grad is None:
                    self.flat[off:off + k].zero_()
                elif p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                    self.flat[off:off + k].copy_(p.grad.reshape(-1))
                off += k
            self.attach()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and self.flat.numel():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, bucket: Optional[torch.Tensor] = None):
    """Sum the .grad of every parameter over all ranks with a single flat all-reduce (for callers that keep their own
    .grad tensors; Trainer uses a GradientArena and reduces in place).  Returns the bucket (reusable)."""
    ps: List[torch.Tensor] = [p for p in params if p.grad is not None]
    if not ps or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bucket
    n = sum(p.grad.numel() for p in ps)
    if bucket is None or bucket.numel() != n or bucket.device != ps[0].grad.device:
        bucket = torch.empty(n, dtype=torch.float32, device=ps[0].grad.device)
    off = 0
    for p in ps:
        k = p.grad.numel()
        bucket[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in ps:
        k = p.grad.numel()
        p.grad.copy_(bucket[off:off + k].view_as(p.grad))
        off += k
    return bucket


class Trainer:
    """forward -> fused masked BCE (+ gradient) -> device BPTT -> in-place gradient all-reduce -> fused Adam."""

    def __init__(self, model: torch.nn.Module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, loss_eps=1e-8, group=None):
        from packages.models._engine import FusedAdam

        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.arena = GradientArena(self.params)
        self.opt = FusedAdam(self.params, lr=lr, betas=betas, eps=eps)
        self.loss_eps = loss_eps
        self.group = group

    def step(self, inputs, target, lengths):
        """inputs: tuple of model inputs (e.g. (audio, video)); returns the local loss (0-dim tensor)."""
        from . import engine as E

        self.model.train()
        logits = self.model(*inputs, lengths)
        loss, _, dlogits = E.batch_bce(logits, target, lengths, self.loss_eps, want_grad=True)
        logits.backward(dlogits)           # autograd accumulates into the arena views
        self.arena.all_reduce(self.group)
        self.opt.step()
        self.arena.zero()
        return loss
