"""Per-device CUDA engines behind the reference-compatible nn.Modules.

The modules in this package keep the reference's parameters/buffers (so ``state_dict`` keys,
``.to()``, ``DataParallel`` replication and checkpoint loading behave as in
sp-uhh/audio-visual-vad) but their ``forward`` bodies call libavvad through ``avvad.engine``.
Packed (BN-folded, bf16, gate-interleaved) weights are cached per device and re-packed whenever a
parameter's storage or version counter changes.
"""
from __future__ import annotations

import os
import sys

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_ROOT not in sys.path:  # make `import avvad` work when only `packages` is on sys.path
    sys.path.insert(0, _PKG_ROOT)

from avvad import engine as E  # noqa: E402
from avvad import lib as L  # noqa: E402


class EngineCache:
    def __init__(self):
        self.by_device = {}

    @staticmethod
    def _signature(module: torch.nn.Module):
        sig = []
        for t in list(module.parameters()) + list(module.buffers()):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def get(self, module: torch.nn.Module, device: torch.device, builder):
        sig = self._signature(module)
        key = (device.type, device.index)
        hit = self.by_device.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        eng = builder(hit[1] if hit is not None else None)
        self.by_device[key] = (sig, eng)
        return eng


def check_inference_only(module: torch.nn.Module):
    if module.training:
        raise NotImplementedError(
            "libavvad implements the eval-mode forward (BatchNorm running statistics) of this module; "
            "call .eval() first.  The training step (batch-statistics BN, backward, Adam) is not built yet.")


def device_of(*tensors) -> torch.device:
    L.require_cuda(*tensors)
    return tensors[0].device
