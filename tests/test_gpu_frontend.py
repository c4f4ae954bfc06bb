"""GPU parity: fused STFT/log-power front end and the upsampling gather, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from oracle import video as ov
from avvad import engine as E
from util import golden, err_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gfe():
    return golden("golden_frontend_34M.npz")


def _wave(g, utt):
    return g[utt + "_wav"].astype(np.float32) / 32768.0


@pytest.mark.parametrize("utt", ["sa1", "sa2", "si494"])
def test_ibm_mask_from_cuda_stft_is_bit_exact(gfe, utt):
    """clean_speech_IBM on the CUDA STFT reproduces the reference's shipped label file exactly."""
    x = ofe.peak_normalise(_wave(gfe, utt))
    n = len(x)
    T = E.stft_num_frames(n)
    assert T == ofe.num_frames(n)
    S = E.stft(torch.tensor(x, device="cuda")[None], [n], [T], T)[0].cpu().numpy()  # (513,T,2)
    mask = ofe.clean_speech_IBM(S[..., 0] + 1j * S[..., 1])
    shape = tuple(gfe[utt + "_ibm_shape"])
    gold = np.unpackbits(gfe[utt + "_ibm_bits"])[: shape[0] * shape[1]].reshape(shape)
    assert mask.shape == shape
    assert int((mask != gold).sum()) == 0


def test_logpower_error_profile_vs_f64_oracle(gfe):
    """Same error profile against the float64 oracle as the fp32 library STFT the reference calls
    (tests/test_oracle_golden.py::test_logpower_f32_vs_f64): median < 5e-6, 99 % < 1e-4, max < 2e-2."""
    mean, std = gfe["audio_mean"].ravel(), gfe["audio_std"].ravel()
    names = ["sa1_noisy", "sa1", "sa2", "si494"]
    waves = [_wave(gfe, u) for u in names]
    nmax = max(len(w) for w in waves)
    batch = np.zeros((len(waves), nmax), np.float32)
    for i, w in enumerate(waves):
        batch[i, : len(w)] = w
    ns = [len(w) for w in waves]
    nf = [ofe.num_frames(n) for n in ns]
    tmax = max(nf) + 3  # force padded rows
    out = E.frontend_logpower(torch.tensor(batch, device="cuda"), ns, nf, tmax, torch.tensor(mean), torch.tensor(std),
                              eps=1e-8, normalise=True).cpu().numpy()
    for i, w in enumerate(waves):
        ref = ofe.frontend_features(w, mean, std, dtype=np.float64)  # (T,513) standardised
        got = out[i, : nf[i]]
        d = np.abs(got - ref) * (std[None, :] + 1e-8)  # back to log-power units
        assert np.quantile(d, 0.5) < 5e-6, err_stats(got, ref)
        assert np.quantile(d, 0.99) < 1e-4, err_stats(got, ref)
        assert d.max() < 2e-2, err_stats(got, ref)
        pad = out[i, nf[i]:]
        expect = ((0.0 - mean) / (std + np.float32(1e-8))).astype(np.float32)
        assert np.array_equal(pad, np.broadcast_to(expect, pad.shape))


def test_frontend_options_and_edges(gfe):
    w = _wave(gfe, "sa2")[:20000]
    n = len(w)
    T = ofe.num_frames(n)
    x = torch.tensor(w, device="cuda")[None]
    raw = E.frontend_logpower(x, [n], [T], T, None, None, normalise=False).cpu().numpy()[0]
    ref = ofe.logpower(w.astype(np.float64), dtype=np.float64).T
    assert np.quantile(np.abs(raw - ref), 0.99) < 1e-4
    # trimming to fewer frames than the STFT has (data_handling.py:483-486)
    trimmed = E.frontend_logpower(x, [n], [T - 5], T - 5, None, None, normalise=False).cpu().numpy()[0]
    assert np.array_equal(trimmed, raw[: T - 5])
    # odd frame counts / single frame
    one = E.frontend_logpower(x[:, :1024], [1024], [1], 1, None, None, normalise=False).cpu().numpy()[0]
    assert np.allclose(one[0], ofe.logpower(w[:1024].astype(np.float64), pad_at_end=False).T[0], atol=1e-3)


def test_frame_count_rule_matches_oracle_everywhere():
    for n in list(range(1024, 1024 + 3000, 7)) + [64000, 73045, 81920, 102741, 160000, 160001]:
        assert E.stft_num_frames(n) == ofe.num_frames(n), n


# ---- upsampling ------------------------------------------------------------------------------------
def test_upsample_index_bit_exact_vs_reference_files():
    g = golden("golden_upsample.npz")
    for tag in g["names"]:
        F, T = int(g[tag + "_F"]), int(g[tag + "_T"])
        idx = E.upsample_index(F, T).cpu().numpy()
        assert np.array_equal(idx, g[tag + "_src"]), tag
        assert E.upsampled_length(F) == ov.upsampled_length(F)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_upsample_gather_bit_exact(dtype):
    rng = np.random.default_rng(0)
    Fs = [152, 131, 7]
    fmax = max(Fs)
    src = rng.integers(0, 256, size=(3, fmax, 67, 67)).astype(np.uint8)
    n_out = [ov.upsampled_length(152), 270, ov.upsampled_length(7) - 1]
    tmax = 320
    t = torch.tensor(src, device="cuda").to(dtype)
    out = E.upsample_gather(t, Fs, n_out, tmax, mean=153.435, std=48.071, eps=1e-8, standardise=True).cpu().numpy()
    for b, F in enumerate(Fs):
        ref = ov.upsample_gather(src[b, :F], n_out[b], 153.435, 48.071, 1e-8)
        assert np.array_equal(out[b, : n_out[b]], ref), b
        pad = (np.float32(0) - np.float32(153.435)) / (np.float32(48.071) + np.float32(1e-8))
        assert np.all(out[b, n_out[b]:] == pad)
    raw = E.upsample_gather(t, Fs, n_out, tmax, standardise=False).cpu().numpy()
    assert np.array_equal(raw[0, :317], src[0, ov.upsample_index(152)].astype(np.float32))
