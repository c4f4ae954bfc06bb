"""Seeded synthetic weights and inputs (SURVEY §8d "Synthetic inputs").

There is no network for datasets or checkpoints, so benchmarks and parity tests run on
random-init weights of the reference architectures and on synthetic audio/video of the
reference's shapes (16 kHz, 1024/256 STFT, 67x67 ROI at 30 fps).  Everything is derived from
(seed, key) so that the reference modules, the oracle and the CUDA path can be loaded with
bit-identical tensors without shipping a checkpoint.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

_STAGES = ((4, 64, 64), (5, 64, 128), (6, 128, 256), (7, 256, 512))


def resnet18_trunk_spec(prefix="features.") -> "OrderedDict[str, Tuple[tuple, torch.dtype]]":
    """Keys/shapes of torchvision resnet18 children[:-1] inside nn.Sequential (AV_Net.py:28-30)."""
    spec: "OrderedDict[str, Tuple[tuple, torch.dtype]]" = OrderedDict()

    def bn(p, c):
        spec[p + ".weight"] = ((c,), torch.float32)
        spec[p + ".bias"] = ((c,), torch.float32)
        spec[p + ".running_mean"] = ((c,), torch.float32)
        spec[p + ".running_var"] = ((c,), torch.float32)
        spec[p + ".num_batches_tracked"] = ((), torch.int64)

    spec[prefix + "0.weight"] = ((64, 3, 7, 7), torch.float32)
    bn(prefix + "1", 64)
    for stage, cin, cout in _STAGES:
        for blk in (0, 1):
            p = f"{prefix}{stage}.{blk}"
            c_in = cin if blk == 0 else cout
            spec[p + ".conv1.weight"] = ((cout, c_in, 3, 3), torch.float32)
            bn(p + ".bn1", cout)
            spec[p + ".conv2.weight"] = ((cout, cout, 3, 3), torch.float32)
            bn(p + ".bn2", cout)
            if blk == 0 and cin != cout:
                spec[p + ".downsample.0.weight"] = ((cout, cin, 1, 1), torch.float32)
                bn(p + ".downsample.1", cout)
    return spec


def lstm_spec(prefix: str, input_size: int, hidden: int, layers: int):
    spec = OrderedDict()
    for l in range(layers):
        i = input_size if l == 0 else hidden
        spec[f"{prefix}.weight_ih_l{l}"] = ((4 * hidden, i), torch.float32)
        spec[f"{prefix}.weight_hh_l{l}"] = ((4 * hidden, hidden), torch.float32)
        spec[f"{prefix}.bias_ih_l{l}"] = ((4 * hidden,), torch.float32)
        spec[f"{prefix}.bias_hh_l{l}"] = ((4 * hidden,), torch.float32)
    return spec


def model_spec(kind: str, y_dim=1, hidden=1024, layers=2, use_mcb=False):
    """state_dict layout of DeepVAD_{AV,audio,video}."""
    spec = OrderedDict()
    if kind == "audio":
        spec.update(lstm_spec("lstm_audio", 513, hidden, layers))
        spec["vad_audio.weight"] = ((y_dim, hidden), torch.float32)
        spec["vad_audio.bias"] = ((y_dim,), torch.float32)
    elif kind == "video":
        spec.update(resnet18_trunk_spec())
        spec.update(lstm_spec("lstm_video", 512, hidden, layers))
        spec["vad_video.weight"] = ((y_dim, hidden), torch.float32)
        spec["vad_video.bias"] = ((y_dim,), torch.float32)
    elif kind == "av":
        spec.update(resnet18_trunk_spec())
        for k in ("weight", "bias", "running_mean", "running_var"):
            spec["bn." + k] = ((512,), torch.float32)
        spec["bn.num_batches_tracked"] = ((), torch.int64)
        if use_mcb:
            spec["mcb.sketch1.h"] = ((513,), torch.int64)
            spec["mcb.sketch1.s"] = ((513,), torch.float32)
            spec["mcb.sketch2.h"] = ((512,), torch.int64)
            spec["mcb.sketch2.s"] = ((512,), torch.float32)
            for k in ("weight", "bias", "running_mean", "running_var"):
                spec["mcb_bn." + k] = ((1024,), torch.float32)
            spec["mcb_bn.num_batches_tracked"] = ((), torch.int64)
            spec.update(lstm_spec("lstm_merged", 1024, hidden, layers))
        else:
            spec.update(lstm_spec("lstm_merged", 1025, hidden, layers))
        spec["vad_merged.weight"] = ((y_dim, hidden), torch.float32)
        spec["vad_merged.bias"] = ((y_dim,), torch.float32)
    else:
        raise ValueError(kind)
    return spec


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def seeded_tensor(key: str, shape, dtype, seed: int) -> torch.Tensor:
    """One tensor of a synthetic state_dict, a pure function of (seed, key, shape)."""
    g = _gen(seed, key)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.int64)
    if leaf == "h":
        return torch.randint(0, 1024, shape, generator=g, dtype=torch.int64)
    if leaf == "s":
        return (2 * torch.randint(0, 2, shape, generator=g) - 1).to(torch.float32)
    if leaf == "running_mean":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "running_var":
        return 0.5 + torch.rand(shape, generator=g)
    is_bn = (".bn" in key or key.startswith("bn.") or key.startswith("mcb_bn.") or ".downsample.1" in key
             or key.startswith("features.1."))
    if is_bn and leaf == "weight":
        return 0.5 + torch.rand(shape, generator=g)
    if is_bn and leaf == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if len(shape) == 4:  # conv: He init so activations keep O(1) scale through the trunk
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    if len(shape) == 3:  # conv1d (WaveNet)
        fan_in = shape[1] * shape[2]
        return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
    if "lstm" in key or key.startswith("vad_"):
        k = 1.0 / 32.0  # PyTorch default U(-1/sqrt(H), 1/sqrt(H)) with H=1024
        return (torch.rand(shape, generator=g) * 2 - 1) * k
    return torch.randn(shape, generator=g) * 0.05


def seeded_state_dict(spec, seed=0) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, seeded_tensor(k, shp, dt, seed)) for k, (shp, dt) in spec.items())


def fill_module_(module: torch.nn.Module, seed=0):
    """Overwrite every tensor in module.state_dict() with its seeded value (in place)."""
    sd = module.state_dict()
    new = OrderedDict((k, seeded_tensor(k, tuple(v.shape), v.dtype, seed)) for k, v in sd.items())
    module.load_state_dict(new)
    return module


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------
def synth_wave(n_samples: int, seed: int) -> np.ndarray:
    """0.1*N(0,1) noise with a 0.5-4 Hz amplitude envelope, clipped to [-1,1]; fp32."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    f_env = rng.uniform(0.5, 4.0)
    env = 0.55 + 0.45 * np.sin(2 * np.pi * f_env * t + rng.uniform(0, 2 * np.pi))
    x = 0.1 * rng.standard_normal(n_samples) * env
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def synth_video_u8(n_frames: int, seed: int, h=67, w=67) -> np.ndarray:
    """Low-pass filtered uniform noise, uint8 (F,67,67)."""
    rng = np.random.default_rng(seed + 7919)
    base = rng.uniform(0, 255, size=(n_frames, h + 4, w + 4))
    k = np.ones(5) / 5.0
    sm = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, base)
    sm = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 2, sm)
    sm = (sm - sm.min()) / (sm.max() - sm.min()) * 255.0
    return np.clip(np.rint(sm), 0, 255).astype(np.uint8)


def synth_audio_stats(seed=0):
    """Per-bin log-power mean/std of the reference's magnitude: N(-2.9,1) / U(1.3,2.4) (513,)."""
    rng = np.random.default_rng(seed + 104729)
    mean = (-2.9 + rng.standard_normal(513)).astype(np.float32)
    std = rng.uniform(1.3, 2.4, 513).astype(np.float32)
    return mean, std


VIDEO_MEAN, VIDEO_STD = 153.435, 48.071  # data/subset/.../matlab_raw/ntcd_timit_statistics.h5
