"""CPU: libavvad.so loads without a GPU, exports every function include/avvad.h declares, the ctypes table covers
them all, and the host-only entry points (frame-count rule, upsampled length) agree with the oracle."""
import ctypes
import os
import re

import pytest

from avvad import lib as L
from oracle import frontend as ofe
from oracle import video as ov

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, "include", "avvad.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avvad_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 30
    l = ctypes.CDLL(L.LIB_PATH)
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, missing
    not_bound = [n for n in names if n not in L.PROTOTYPES]
    assert not not_bound, not_bound


def test_host_only_entry_points_match_oracle():
    l = L.lib()
    assert l.avvad_version() >= 100
    for n in (1024, 1279, 1280, 64000, 73045, 81920, 102741):
        assert l.avvad_stft_num_frames(n, 16000.0, 0.064, 0.25, 1) == ofe.num_frames(n)
    for f in (1, 6, 18, 131, 152, 192):
        assert l.avvad_upsampled_length(f, 25, 12) == ov.upsampled_length(f)


def test_compute_entry_points_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from packages.models.Audio_Net import DeepVAD_audio
    m = DeepVAD_audio(1, 64, 1).eval()
    with pytest.raises(L.AvvadError):
        m(torch.zeros(1, 2, 513), [2])
