"""Oracle: audio front end and label generation.  Test infrastructure only.

Follows the reference:
  * pad-at-end rule, periodic Hann, framing ........ packages/processing/stft.py:123-151
  * peak normalisation .............................. packages/data_handling.py:441
  * power / log ..................................... packages/data_handling.py:454-457
  * standardisation ................................. scripts/evaluate_AV_net.py:228-230
  * clean_speech_VAD / clean_speech_IBM ............. packages/processing/target.py:5-70
"""
from __future__ import annotations

import math

import numpy as np


def stft_params(fs=16000, wlen_sec=64e-3, hop_percent=0.25):
    """stft.py:123-126 -- nfft=int(wlen_sec*fs), hop=int(hop_percent*nfft)."""
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of STFT is not an integer.")
    nfft = int(wlen_sec * fs)
    hop = int(hop_percent * nfft)
    return nfft, hop


def pad_at_end_fires(n_samples: int, fs=16000, wlen_sec=64e-3, hop_percent=0.25) -> bool:
    """stft.py:134-139 -- same double-precision expression, evaluated in the same order."""
    utt_len = n_samples / fs
    q = utt_len / wlen_sec / hop_percent
    return math.ceil(q) != int(q)


def num_frames(n_samples: int, fs=16000, wlen_sec=64e-3, hop_percent=0.25, pad_at_end=True) -> int:
    """torch.stft(center=False): 1 + (N' - nfft)//hop with N' the padded length."""
    nfft, hop = stft_params(fs, wlen_sec, hop_percent)
    n = n_samples + (hop if (pad_at_end and pad_at_end_fires(n_samples, fs, wlen_sec, hop_percent)) else 0)
    if n < nfft:
        return 0
    return 1 + (n - nfft) // hop


def hann_periodic(nfft: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(nfft) (periodic=True): 0.5 - 0.5 cos(2 pi n / nfft)."""
    n = np.arange(nfft, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / nfft)).astype(dtype)


def frame_signal(x: np.ndarray, nfft: int, hop: int) -> np.ndarray:
    """(T, nfft) view of hop-strided frames (center=False)."""
    T = 1 + (len(x) - nfft) // hop
    return np.lib.stride_tricks.sliding_window_view(x, nfft)[::hop][:T]


def stft_complex(x: np.ndarray, fs=16000, wlen_sec=64e-3, hop_percent=0.25, pad_at_end=True,
                 dtype=np.float64) -> np.ndarray:
    """stft_pytorch(center=False) -> complex (F, T).  dtype float64 = high-precision oracle,
    float32 = what the reference's fp32 torch.stft computes (up to FFT rounding)."""
    nfft, hop = stft_params(fs, wlen_sec, hop_percent)
    x = np.asarray(x, dtype=dtype)
    if pad_at_end and pad_at_end_fires(len(x), fs, wlen_sec, hop_percent):
        x = np.concatenate([x, np.zeros(hop, dtype=dtype)])
    frames = frame_signal(x, nfft, hop) * hann_periodic(nfft, dtype)[None, :]
    spec = np.fft.rfft(frames, axis=1)
    return spec.T  # (F, T)


def stft_torch_fp32(x: np.ndarray, fs=16000, wlen_sec=64e-3, hop_percent=0.25, pad_at_end=True):
    """Exactly the reference's library call (stft.py:145-151) with the modern return_complex API."""
    import torch

    nfft, hop = stft_params(fs, wlen_sec, hop_percent)
    xt = torch.as_tensor(np.asarray(x, dtype=np.float32))
    if pad_at_end and pad_at_end_fires(len(xt), fs, wlen_sec, hop_percent):
        xt = torch.nn.functional.pad(xt, (0, hop), mode="constant")
    S = torch.stft(xt, n_fft=nfft, hop_length=hop, win_length=None, window=torch.hann_window(nfft),
                   center=False, return_complex=True)
    return S.numpy()


def peak_normalise(x: np.ndarray) -> np.ndarray:
    """data_handling.py:441 -- x / max|x| in fp32."""
    x = np.asarray(x, dtype=np.float32)
    return x / np.max(np.abs(x))


def logpower(x: np.ndarray, eps=1e-8, dtype=np.float64, **kw) -> np.ndarray:
    """data_handling.py:454-457 -- log(re^2 + im^2 + eps) -> (F, T)."""
    S = stft_complex(x, dtype=dtype, **kw)
    p = S.real ** 2 + S.imag ** 2
    return np.log(p + dtype(eps))


def standardise(feat_ft: np.ndarray, mean: np.ndarray, std: np.ndarray, eps=1e-8) -> np.ndarray:
    """evaluate_AV_net.py:225-230 -- (F,T)->(T,F); (x - mean.T)/(std+eps).T."""
    x = feat_ft.T
    return (x - mean.reshape(1, -1)) / (std.reshape(1, -1) + eps)


def frontend_features(wave: np.ndarray, mean, std, eps=1e-8, normalise=True, dtype=np.float64,
                      n_frames=None) -> np.ndarray:
    """A3 -> A1 -> A2 -> (trim) -> A4.  Returns (T, 513)."""
    x = peak_normalise(wave) if normalise else np.asarray(wave, dtype=np.float32)
    lp = logpower(x.astype(dtype), eps=eps, dtype=dtype)
    if n_frames is not None:
        lp = lp[:, :n_frames]
    return standardise(lp, np.asarray(mean, dtype=dtype), np.asarray(std, dtype=dtype), eps).astype(dtype)


# ---------------------------------------------------------------------------------------------
# labels (target.py)
# ---------------------------------------------------------------------------------------------
def clean_speech_VAD(speech_t: np.ndarray, fs=16000, wlen_sec=64e-3, hop_percent=0.25, center=False,
                     pad_at_end=True, vad_threshold=1.70) -> np.ndarray:
    """target.py:5-56 -- frame energy > 10**thr * min(frame energy); librosa.util.frame replaced
    by a strided view.  Returns float32 (1, T)."""
    nfft, hop = stft_params(fs, wlen_sec, hop_percent)
    y = np.asarray(speech_t)
    if pad_at_end and pad_at_end_fires(len(y), fs, wlen_sec, hop_percent):
        y = np.pad(y, (0, hop), mode="constant")
    if center:
        y = np.pad(y, int(nfft // 2), mode="reflect")
    frames = frame_signal(y, nfft, hop).T  # (nfft, T) like librosa.util.frame
    power = np.power(frames, 2).sum(axis=0)
    vad = power > np.power(10, vad_threshold) * np.min(power)
    return np.float32(vad)[None]


def clean_speech_IBM(speech_tf: np.ndarray, eps=1e-8, ibm_threshold=50) -> np.ndarray:
    """target.py:58-70 -- 20 log10(|S|+eps) > max - thr."""
    mag = abs(speech_tf)
    power_db = 20 * np.log10(mag + eps)
    mask = power_db > np.max(power_db) - ibm_threshold
    return np.float32(mask)
