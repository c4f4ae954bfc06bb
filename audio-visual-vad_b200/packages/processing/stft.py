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
    # the input's own device when it is a CUDA tensor, otherwise the calling thread's current CUDA device
    dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        xd = x.detach().to(device=dev, dtype=torch.float32)
        out = _E.stft(xd[None], [n], [T], T)[0]  # (513, T, 2)
    return out.to(x.device)


def stft_frame_count(n_samples, fs=16e3, wlen_sec=64e-3, hop_percent=0.25, pad_at_end=True):
    """Number of frames stft_pytorch(center=False) returns for a signal of n_samples (host arithmetic only: safe in
    forked DataLoader workers)."""
    return _E.stft_num_frames(int(n_samples), fs, wlen_sec, hop_percent, pad_at_end)


# ---- host-side numpy STFT / inverse STFT (packages/processing/stft.py:13-99) --------------------------------------
# The reference wraps librosa.core.stft / istft; these are the same transforms written out in numpy (periodic window,
# optional centre padding, window-sum-square normalised overlap-add).  They serve the reconstruction and plotting
# scripts (scripts/reconstruct_dnn_classif.py, scripts/visualization_*.py), which run on the host in the reference too;
# the training / evaluation hot path uses stft_pytorch above.

def _frame_params(fs, wlen_sec, hop_percent, what):
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError(f"wlen_sample of {what} is not an integer.")
    nfft = int(wlen_sec * fs)
    return nfft, int(hop_percent * nfft)


def _window(win, nfft):
    import numpy as np
    if isinstance(win, str):
        from scipy.signal import get_window
        return get_window(win, nfft, fftbins=True)
    w = np.asarray(win, dtype=np.float64)
    if w.shape != (nfft,):
        raise ValueError("window must have nfft samples")
    return w


def stft(x, fs=16e3, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, pad_mode='reflect', pad_at_end=True,
         dtype='complex64'):
    """x: 1-D float array.  Returns the complex spectrogram (1 + nfft/2, T)."""
    import math
    import numpy as np
    nfft, hop = _frame_params(fs, wlen_sec, hop_percent, "STFT")
    x = np.asarray(x)
    if pad_at_end:
        utt_len = len(x) / fs
        if math.ceil(utt_len / wlen_sec / hop_percent) != int(utt_len / wlen_sec / hop_percent):
            x = np.pad(x, (0, hop), mode='constant')
    if center:
        x = np.pad(x, nfft // 2, mode=pad_mode)
    if len(x) < nfft:
        raise ValueError("signal shorter than one STFT window")
    n_frames = 1 + (len(x) - nfft) // hop
    x = np.ascontiguousarray(x, dtype=np.float64)
    frames = np.lib.stride_tricks.as_strided(x, shape=(n_frames, nfft), strides=(hop * x.strides[0], x.strides[0]))
    spec = np.fft.rfft(frames * _window(win, nfft), axis=1)
    return np.ascontiguousarray(spec.T).astype(dtype)


def istft(Sxx, fs=16000, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, dtype='float32', max_len=None):
    """Sxx: (1 + nfft/2, T) complex.  Overlap-add inverse of stft; `max_len` is the wanted length in samples."""
    import numpy as np
    nfft, hop = _frame_params(fs, wlen_sec, hop_percent, "iSTFT")
    Sxx = np.asarray(Sxx)
    n_frames = Sxx.shape[1]
    w = _window(win, nfft)
    frames = np.fft.irfft(Sxx.T, n=nfft, axis=1) * w
    total = nfft + hop * (n_frames - 1)
    y = np.zeros(total)
    wss = np.zeros(total)
    for t in range(n_frames):
        y[t * hop:t * hop + nfft] += frames[t]
        wss[t * hop:t * hop + nfft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    start = nfft // 2 if center else 0
    if max_len is None:
        y = y[start:total - start] if center else y
    else:
        y = y[start:]
        y = y[:max_len] if len(y) >= max_len else np.pad(y, (0, max_len - len(y)))
        y = y[:int(max_len * fs)]  # the reference's extra trim (stft.py:98), a no-op for lengths in samples
    return y.astype(dtype)
