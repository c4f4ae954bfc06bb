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


# ---- threshold-based masks (packages/processing/target.py:110-251; not used by the VAD scripts) --------------------

def _raised_cosine(width):
    """(1 + cos(pi k / (width - 1))) / 2, k = 0 .. width-1: a falling edge from 1 to 0 (angle formed as
    (pi / (width - 1)) * k, the reference's order of operations, so the tables agree to the last bit)."""
    return np.cos(np.pi / (width - 1) * np.arange(0, width))


def _voiced_unvoiced_split_characteristic(number_of_frequency_bins):
    """Two complementary frequency weightings: `voiced` passes bins 4..~250 (fast rise at the bottom, 99-bin
    raised-cosine roll-off centred on bin 200), `unvoiced` is its mirror above the split, cut off at bin 500."""
    split_bin, width, fast, low_bin, high_bin = 200, 99, 5, 4, 500
    edge, fast_edge = 0.5 * (1 + _raised_cosine(width)), (_raised_cosine(fast) + 1) / 2
    start = int(split_bin - width / 2)
    voiced = np.ones(number_of_frequency_bins)
    voiced[start - 1:start + width - 1] = edge
    voiced[start - 1 + width:] = 0
    voiced[:low_bin] = 0
    voiced[low_bin - 1:low_bin + fast - 1] = 1 - fast_edge
    unvoiced = np.ones(number_of_frequency_bins)
    unvoiced[start - 1:start + width - 1] = 1 - edge
    unvoiced[:start] = 0
    unvoiced[high_bin - 1:] = 0
    unvoiced[high_bin - 1:high_bin + fast - 1] = fast_edge
    return voiced, unvoiced


def _thresholded_psd(X, t_voiced, t_unvoiced):
    voiced, unvoiced = _voiced_unvoiced_split_characteristic(X.shape[-1])
    threshold_db = t_voiced * voiced + t_unvoiced * unvoiced
    return (X * X.conjugate()) / np.power(10, threshold_db / 10)


def noise_aware_IBM(X, N, threshold_unvoiced_speech=5, threshold_voiced_speech=0, threshold_unvoiced_noise=-10,
                    threshold_voiced_noise=-10, low_cut=5, high_cut=500):
    """(speech mask, noise mask) for STFTs X, N of shape (frames, bins): speech where the thresholded speech power
    exceeds the noise power (and 0.005); noise where a second threshold falls below it; bins outside
    [low_cut - 1, high_cut) are forced to 0 / 1."""
    x_speech = _thresholded_psd(X, threshold_voiced_speech, threshold_unvoiced_speech)
    x_noise = _thresholded_psd(X, threshold_unvoiced_noise, threshold_voiced_noise)  # argument order as in the reference
    n_psd = N * N.conjugate()
    speech = np.logical_and(x_speech > n_psd, x_speech > 0.005)
    speech[..., 0:low_cut - 1] = 0
    speech[..., high_cut:len(speech[0])] = 0
    noise = np.logical_or(x_noise < n_psd, x_noise < 0.005)
    noise[..., 0:low_cut - 1] = 1
    noise[..., high_cut:len(noise[0])] = 1
    return speech, noise


def threshold_IBM(X, threshold_unvoiced_speech=5, threshold_voiced_speech=0, threshold_unvoiced_noise=-10,
                  threshold_voiced_noise=-10, low_cut=5, high_cut=500):
    """Speech mask of noise_aware_IBM against a constant noise power of 10."""
    x_speech = _thresholded_psd(X, threshold_voiced_speech, threshold_unvoiced_speech)
    speech = np.logical_and(x_speech > 10, x_speech > 0.005)
    speech[..., 0:low_cut - 1] = 0
    speech[..., high_cut:len(speech[0])] = 0
    return speech
