"""GPU parity: WaveNet encoder (dead code in the reference but named by north_star) vs the reference's own output."""
import numpy as np
import pytest
import torch

from avvad import synth
from oracle import models as om
from util import golden, err_stats

pytestmark = pytest.mark.gpu


def test_wavenet_encoder_matches_reference_golden():
    from packages.models.wavenet_autoencoder import wavenet_autoencoder
    g = golden("ref_models.npz")
    wn = wavenet_autoencoder(filter_width=2, quantization_channel=16, dilations=[1, 2, 4, 8, 1, 2, 4, 8],
                             en_residual_channel=32, en_dilation_channel=32, en_bottleneck_width=16,
                             en_pool_kernel_size=10, use_bias=True)
    synth.fill_module_(wn, seed=16)
    wn = wn.cuda().eval()
    out = wn(torch.tensor(g["wavenet_x"]).cuda()).cpu().numpy()
    st = err_stats(out, g["wavenet_out"])
    assert out.shape == g["wavenet_out"].shape
    assert st["rel_fro"] < 2e-2 and st["max"] < 3e-2 * max(1.0, st["ref_absmax"]), st


def test_wavenet_wider_config_vs_oracle():
    from packages.models.wavenet_autoencoder import wavenet_autoencoder
    dil = [1, 2, 4, 8, 16, 32]
    wn = wavenet_autoencoder(3, 64, dil, 128, 64, 48, 7, use_bias=False)
    synth.fill_module_(wn, seed=5)
    x = torch.randn(3, 64, 700, generator=torch.Generator().manual_seed(1))
    sd = {k: v for k, v in wn.state_dict().items()}
    ref = om.wavenet_encode(x, sd, dil, 7).numpy()
    out = wn.cuda().eval()(x.cuda()).cpu().numpy()
    st = err_stats(out, ref)
    assert st["rel_fro"] < 2e-2, st
