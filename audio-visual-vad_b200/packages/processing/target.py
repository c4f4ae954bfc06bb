"""Label generation with the reference's signatures (packages/processing/target.py:5-104): energy-threshold VAD on
framed clean speech and the ideal binary mask.  Host-side numpy (run once per dataset); librosa.util.frame is replaced
by a strided view.  These reproduce the reference's shipped label files bit-for-bit (tests/test_oracle_golden.py)."""
import math

import numpy as np


def _frames(y, nfft, hop):
    n = 1 + (len(y) - nfft) // hop
    return np.lib.stride_tricks.sliding_window_view(y, nfft)[::hop][:n].T  # (nfft, T) like librosa.util.frame


def clean_speech_VAD(speech_t, fs=16e3, wlen_sec=50e-3, hop_percent=0.25, center=True, pad_mode='reflect',
                     pad_at_end=True, vad_threshold=1.70):
    nfft = int(wlen_sec * fs)
    hopsamp = int(hop_percent * nfft)
    y = np.asarray(speech_t)
    if pad_at_end:
        utt_len = len(y) / fs
        if math.ceil(utt_len / wlen_sec / hop_percent) != int(utt_len / wlen_sec / hop_percent):
            y = np.pad(y, (0, hopsamp), mode='constant')
    if center:
        y = np.pad(y, int(nfft // 2), mode=pad_mode)
    power = np.power(_frames(y, nfft, hopsamp), 2).sum(axis=0)
    vad = power > np.power(10, vad_threshold) * np.min(power)
    return np.float32(vad)[None]


def clean_speech_IBM(speech_tf, eps=1e-8, ibm_threshold=50):
    power_db = 20 * np.log10(abs(speech_tf) + eps)
    return np.float32(power_db > np.max(power_db) - ibm_threshold)


def noise_robust_clean_speech_IBM(speech_t, speech_tf, fs=16e3, wlen_sec=50e-3, hop_percent=0.25, center=True,
                                  pad_mode='reflect', pad_at_end=True, vad_threshold=1.70, eps=1e-8, ibm_threshold=50):
    vad = clean_speech_VAD(speech_t, fs=fs, wlen_sec=wlen_sec, hop_percent=hop_percent, center=center,
                           pad_mode=pad_mode, pad_at_end=pad_at_end, vad_threshold=vad_threshold)
    return clean_speech_IBM(speech_tf, eps=eps, ibm_threshold=ibm_threshold) * vad
