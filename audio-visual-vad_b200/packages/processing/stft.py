"""STFT front end with the reference's signature (packages/processing/stft.py:102-152), computed by
the fused sm_100a kernel in libavvad (csrc/frontend.cu).

Only the configuration every reference script uses is implemented on the device: 16 kHz, 64 ms
Hann window (nfft 1024), 25 % hop (256), center=False.  Other configurations raise instead of
silently falling back to a library STFT."""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from avvad import engine as _E  # noqa: E402


def stft_pytorch(x, fs=16e3, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, pad_mode='reflect',
                 pad_at_end=True):
    """x: 1-D float tensor.  Returns the legacy real view (F, T, 2) on x's device."""
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of STFT is not an integer.")
    nfft = int(wlen_sec * fs)
    hop = int(hop_percent * nfft)
    if nfft != 1024 or hop != 256 or center or win != 'hann':
        raise NotImplementedError("libavvad front end supports nfft=1024, hop=256, win='hann', center=False "
                                  f"(got nfft={nfft}, hop={hop}, win={win!r}, center={center})")
    if x.dim() != 1:
        raise ValueError("stft_pytorch expects a 1-D signal")
    n = x.shape[0]
    T = _E.stft_num_frames(n, fs, wlen_sec, hop_percent, pad_at_end)
    xd = x.detach().to(device='cuda', dtype=torch.float32)
    out = _E.stft(xd[None], [n], [T], T)[0]  # (513, T, 2)
    return out.to(x.device)
