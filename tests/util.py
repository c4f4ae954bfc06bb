"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def err_stats(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    return {"max": float(d.max()), "mean": float(d.mean()), "p99": float(np.quantile(d, 0.99)),
            "ref_absmax": float(np.abs(b).max()), "rel_fro": float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))}


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)
