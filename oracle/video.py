"""Oracle: video data path.  Test infrastructure only.

Follows the reference:
  * 30 -> 62.5 fps frame-rate conversion ............ scripts/create_video_train_files_upsampled.py:58-59,116-173
    (done there by ``ffmpeg -filter:v fps=fps=62.5`` + libx264 crf 0; the equivalent index map
    was recovered from, and is pinned by, the reference's own ``*_upsampled.h5`` files --
    tests/golden/upsample_index.npz)
  * DCT -> ROI image ................................. same file :137-162 and packages/processing/video.py:5-24
  * (H,W,T)->(T,H,W), standardise ................... scripts/evaluate_AV_net.py:176-182
"""
from __future__ import annotations

import numpy as np

# 62.5 / 30 as an exact rational (create_video_train_files_upsampled.py:58-59)
FPS_NUM, FPS_DEN = 25, 12


def upsampled_length(n_src: int, num=FPS_NUM, den=FPS_DEN) -> int:
    """T_up = floor(F*num/den + 1/2): number of output slots k whose timestamp is covered."""
    return (2 * n_src * num + den) // (2 * den)


def upsample_index(n_src: int, n_out: int | None = None, num=FPS_NUM, den=FPS_DEN) -> np.ndarray:
    """src(k) = max{i : floor(i*num/den + 1/2) <= k}  (ffmpeg fps filter, round-to-nearest
    timestamps) = (den*(2k+1) - 1) // (2*num);  k in [0, T_up) optionally trimmed to n_out
    (create_video_train_files_upsampled.py:238-241 trims to the label length)."""
    T = upsampled_length(n_src, num, den)
    if n_out is not None:
        T = min(T, int(n_out))
    k = np.arange(T, dtype=np.int64)
    src = (den * (2 * k + 1) - 1) // (2 * num)
    return np.minimum(src, n_src - 1)


def upsample_index_bruteforce(n_src: int, num=FPS_NUM, den=FPS_DEN) -> np.ndarray:
    """Literal statement of the model (used to check the closed form)."""
    start = [(2 * i * num + den) // (2 * den) for i in range(n_src + 1)]
    out = []
    for k in range(start[n_src]):
        i = max(j for j in range(n_src) if start[j] <= k)
        out.append(i)
    return np.asarray(out, dtype=np.int64)


def upsample_gather(frames: np.ndarray, n_out=None, mean=None, std=None, eps=1e-8) -> np.ndarray:
    """(F,H,W) u8/f32 -> (T,H,W) f32; optional (x-mean)/(std+eps) (evaluate_AV_net.py:180-182)."""
    idx = upsample_index(frames.shape[0], n_out)
    out = frames[idx].astype(np.float32)
    if mean is not None:
        out = (out - np.float32(mean)) / (np.float32(std) + np.float32(eps))
    return out


def _idct_matrix(n: int) -> np.ndarray:
    """scipy.fftpack.idct(x) (type 2, norm=None) as a matrix: y[k] = x[0] + 2 sum_{j>=1} x[j] cos(pi (2k+1) j / 2n)."""
    k = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(n, dtype=np.float64)[None, :]
    C = 2.0 * np.cos(np.pi * (2 * k + 1) * j / (2 * n))
    C[:, 0] = 1.0
    return C


def dct_to_roi(mat_row: np.ndarray, width=67, height=67) -> np.ndarray:
    """create_video_train_files_upsampled.py:146-150: idct(idct(reshaped).T).T, float64."""
    a = np.asarray(mat_row, dtype=np.float64).reshape(width, height)
    C = _idct_matrix(height)
    # idct along last axis of a, transpose, idct along last axis, transpose back
    step1 = a @ C.T
    step2 = (step1.T @ _idct_matrix(width).T).T
    return step2


def roi_to_u8_per_frame(idct_frame: np.ndarray) -> np.ndarray:
    """Variant that produced the shipped ``*_upsampled.h5`` goldens (packages/processing/video.py:14,
    the commented cv2.normalize(..., 255, 0, NORM_MINMAX, CV_8U) line) followed by rot90(.,3)
    (video.py:15).  Pinned to +-1 grey level (x264 RGB<->YUV round trip) by tests/golden."""
    lo, hi = idct_frame.min(), idct_frame.max()
    scaled = (idct_frame - lo) * (255.0 / (hi - lo))
    u8 = np.clip(np.rint(scaled), 0, 255).astype(np.uint8)
    return np.rot90(u8, 3)


def roi_to_u8_global(idct_frames: np.ndarray) -> np.ndarray:
    """The script as shipped (create_video_train_files_upsampled.py:140-158): global min and the
    largest per-row (max-min) over all frames; float result is truncated by the uint8 writer."""
    A = idct_frames
    denom = (A.max(axis=-1) - A.min(axis=-1)).max()
    out = (A - A.min()) / denom * 255.0
    return np.stack([np.rot90(f, 3) for f in out])
