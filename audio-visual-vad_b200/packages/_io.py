"""File access for the dataset classes: h5py / torchaudio when they are installed (as in the reference), otherwise the
bundled pure-Python HDF5 reader (avvad/h5min.py: superblock v0, chunked + LZF/deflate -- every file the reference
ships) and the standard-library wave module (16-bit PCM / 32768, which is what torchaudio.load returned)."""
import os
import sys

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def read_h5(path, key):
    try:
        import h5py
    except ImportError:
        from avvad.h5min import H5File
        return H5File(path)[key]
    with h5py.File(path, "r") as f:
        return np.array(f[key][:])


def load_wav(path):
    """(waveform (channels, n) float32 tensor, sample_rate) like torchaudio.load."""
    try:
        import torchaudio
        return torchaudio.load(path)
    except Exception:
        from avvad.h5min import read_wav_int16
        data, fs = read_wav_int16(path)
        return torch.from_numpy(data.astype(np.float32) / 32768.0)[None], fs
