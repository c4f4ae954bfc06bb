"""Data-parallel training step (BASELINE config 5): one process per GPU, local batch B/N, identical replicas,
and ONE all-reduce (sum) of the trainable gradients per step -- what remains of nn.DataParallel's per-step parameter
broadcast / scatter / gather / reduce-add (scripts/train_AV_net.py:193,293-307).

The loss is a SUM over utterances (scripts/train_AV_net.py:298-302), so gradients add across ranks without rescaling.
BatchNorm batch statistics and the MCB whole-tensor L2 norm stay per rank, exactly as they are per DataParallel
replica in the reference.  torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, bucket: Optional[torch.Tensor] = None):
    """Sum the .grad of every parameter over all ranks with a single flat all-reduce.  Returns the bucket (reusable)."""
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
    """forward -> fused masked BCE (+ gradient) -> device BPTT -> gradient all-reduce -> fused Adam."""

    def __init__(self, model: torch.nn.Module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, loss_eps=1e-8, group=None):
        from packages.models._engine import FusedAdam

        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = FusedAdam(self.params, lr=lr, betas=betas, eps=eps)
        self.loss_eps = loss_eps
        self.group = group
        self.bucket = None

    def step(self, inputs, target, lengths):
        """inputs: tuple of model inputs (e.g. (audio, video)); returns the local loss (0-dim tensor)."""
        from . import engine as E

        self.model.train()
        logits = self.model(*inputs, lengths)
        loss, _, dlogits = E.batch_bce(logits, target, lengths, self.loss_eps, want_grad=True)
        logits.backward(dlogits)
        self.bucket = allreduce_gradients(self.params, self.group, self.bucket)
        self.opt.step()
        self.opt.zero_grad(set_to_none=False)
        return loss
