"""DCT -> mouth-ROI image helper (packages/processing/video.py:5-24): 2-D inverse DCT of one frame of the NTCD-TIMIT
.mat coefficients, global min/max normalisation, rot90(.,3), optional label square, 3-channel merge."""
import numpy as np


def _idct_unnormalised(x):
    """scipy.fftpack.idct(x) (type 2, norm=None) along the last axis."""
    n = x.shape[-1]
    k = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    C = 2.0 * np.cos(np.pi * (2 * k + 1) * j / (2 * n))
    C[:, 0] = 1.0
    return x @ C.T


def preprocess_ntcd_matlab(matlab_frames, frame, width, height, y_hat_hard=None, output_video=True):
    df = matlab_frames[frame].reshape(width, height)
    idct_df = _idct_unnormalised(_idct_unnormalised(df).T).T
    A = _idct_unnormalised(_idct_unnormalised(matlab_frames.reshape(-1, width, height)))
    normalized = (idct_df - A.min()) / (A.max(axis=-1) - A.min(axis=-1)).max() * 255.0
    rotated = np.rot90(normalized, 3).copy()
    if y_hat_hard is not None and y_hat_hard[frame] == 1:
        rotated[-9:, -9:] = 255
    return np.stack([rotated] * 3, axis=-1)
