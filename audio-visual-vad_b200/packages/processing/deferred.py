"""Deferred log-power spectrograms: how the UNCHANGED training scripts keep their fork-based DataLoader workers.

The reference's datasets compute the STFT inside ``__getitem__`` (packages/data_handling.py:441-457), i.e. inside the
DataLoader workers that scripts/train_AV_net.py:141-146 forks (num_workers=16, default context) AFTER the parent has put
the model on the GPU.  Here the STFT is a CUDA kernel and a forked child of a CUDA-initialised process cannot touch CUDA,
and there is no CPU implementation of the front end to fall back to.  So a worker only does the file I/O and returns a
:class:`DeferredLogPower` (raw waveform + the number of frames to keep); the collate function (which also runs in the
worker) stacks them into a :class:`DeferredLogPowerBatch`; that object pickles as "call materialise_batch(...) on load",
so the moment the PARENT process takes the batch off the worker queue it runs ONE batched ``avvad_frontend_logpower``
on the parent's current CUDA device and the training loop receives the plain ``(B, T, 513)`` tensor it expects
(on the CPU, because the loaders pin and the scripts ``.to(device)`` what they get).  With num_workers=0 everything
happens in place.  Peak normalisation (data_handling.py:441) is fused into the kernel.
"""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

_materialiser = None   # test hook: (waves (B,N) cpu f32, n_samples, n_frames, t_max, eps) -> (B, t_max, 513) cpu f32


def set_materialiser(fn):
    """Test hook for boxes without a GPU (the CPU suite injects the oracle front end); None restores the CUDA path."""
    global _materialiser
    _materialiser = fn


def in_worker() -> bool:
    from torch.utils.data import get_worker_info
    return get_worker_info() is not None


def materialise_batch(waves, n_samples, n_frames, t_max, eps):
    """(B,N) zero-padded raw waveforms -> (B, t_max, 513) log-power, rows past n_frames[b] zero; CPU tensor."""
    if _materialiser is not None:
        return _materialiser(waves, n_samples, n_frames, t_max, eps)
    from avvad import engine as E
    from avvad import lib as L
    L.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    out = E.frontend_logpower(waves.to(dev, torch.float32, non_blocking=True), n_samples, n_frames, t_max, None, None,
                              eps, True)
    return out.cpu()


class DeferredLogPower:
    """One utterance's log-power spectrogram (513, T), not computed yet.  Supports what the dataset / collate code does
    with the tensor it stands for: ``.shape``, ``.size()``, trimming ``x[..., :n]`` and ``.materialise()``."""

    def __init__(self, wave: torch.Tensor, n_frames: int, eps: float):
        self.wave, self.n_frames, self.eps = wave, int(n_frames), float(eps)

    @property
    def shape(self):
        return torch.Size((513, self.n_frames))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def __getitem__(self, idx):
        if not (isinstance(idx, tuple) and len(idx) == 2 and idx[0] is Ellipsis and isinstance(idx[1], slice)
                and idx[1].start in (None, 0) and idx[1].step in (None, 1)):
            raise IndexError("a deferred spectrogram only supports x[..., :n]")
        n = self.n_frames if idx[1].stop is None else max(0, min(self.n_frames, int(idx[1].stop)))
        return DeferredLogPower(self.wave, n, self.eps)

    def materialise(self) -> torch.Tensor:
        n = self.wave.shape[-1]
        return materialise_batch(self.wave[None], [n], [self.n_frames], self.n_frames, self.eps)[0].t().contiguous()


class DeferredLogPowerBatch:
    """A collated batch of deferred spectrograms; unpickling it (= the parent receiving it from a worker) computes it."""

    def __init__(self, items):
        self.n_samples = [int(it.wave.shape[-1]) for it in items]
        self.n_frames = [it.n_frames for it in items]
        self.eps = items[0].eps
        self.waves = torch.zeros(len(items), max(self.n_samples), dtype=torch.float32)
        for i, it in enumerate(items):
            self.waves[i, : self.n_samples[i]] = it.wave

    def materialise(self, t_max=None) -> torch.Tensor:
        return materialise_batch(self.waves, self.n_samples, self.n_frames, t_max or max(self.n_frames), self.eps)

    def __reduce__(self):
        return materialise_batch, (self.waves, self.n_samples, self.n_frames, max(self.n_frames), self.eps)
