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


def _lib_profile(x_norm, ref64):
    """Error quantiles (log-power units, vs the float64 oracle) of the fp32 library STFT the reference
    itself calls (torch.stft, stft.py:145).  fp32 FFT noise dominates deep spectral nulls of quiet
    frames (up to ~5e-2 on clean speech), so the CUDA kernel is held to this profile, not to a
    fixed absolute number."""
    S = ofe.stft_torch_fp32(x_norm)
    lp = np.log(S.real.astype(np.float32) ** 2 + S.imag.astype(np.float32) ** 2 + np.float32(1e-8)).T
    d = np.abs(lp - ref64)
    return {q: float(np.quantile(d, q)) for q in (0.5, 0.9, 0.99, 0.999, 1.0)}


def _assert_profile(got_lp, ref64, lib, what):
    d = np.abs(got_lp - ref64)
    mine = {q: float(np.quantile(d, q)) for q in lib}
    for q in lib:
        assert mine[q] <= 1.5 * lib[q] + 2e-6, (what, q, mine, lib)


def test_logpower_error_profile_vs_f64_oracle(gfe):
    """Against the float64 oracle the CUDA front end is at most 1.5x as far, at every quantile, as the
    fp32 library STFT the reference calls on the same signal."""
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
        xn = ofe.peak_normalise(w)
        ref = ofe.logpower(xn, dtype=np.float64).T  # (T,513) log-power
        got = out[i, : nf[i]].astype(np.float64) * (std[None, :] + 1e-8) + mean[None, :]  # undo standardisation
        _assert_profile(got, ref, _lib_profile(xn, ref), names[i])
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
    _assert_profile(raw, ref, _lib_profile(w, ref), "sa2[:20000] raw")
    # trimming to fewer frames than the STFT has (data_handling.py:483-486)
    trimmed = E.frontend_logpower(x, [n], [T - 5], T - 5, None, None, normalise=False).cpu().numpy()[0]
    assert np.array_equal(trimmed, raw[: T - 5])
    # odd frame counts / single frame
    seg = np.ascontiguousarray(_wave(gfe, "sa1_noisy")[30000:31024])
    one = E.frontend_logpower(torch.tensor(seg, device="cuda")[None], [1024], [1], 1, None, None,
                              normalise=False).cpu().numpy()[0]
    assert one.shape == (1, 513)
    assert np.allclose(one[0], ofe.logpower(seg.astype(np.float64), pad_at_end=False).T[0], atol=5e-3)


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
