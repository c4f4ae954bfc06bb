"""GPU parity for the data-preparation kernels: DCT -> mouth-ROI decode, VAD / IBM label
generation and dataset statistics, against the numpy oracle and the reference's own shipped files (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from oracle import video as ov
from avvad import engine as E
from util import golden

pytestmark = pytest.mark.gpu


def test_dct_decode_per_frame_matches_oracle_and_reference_pixels():
    gup = golden("golden_upsample.npz")
    rows = gup["sa1_mat_rows"]                                  # (12, 4489) real coefficients of test/34M/sa1
    rng = np.random.default_rng(0)
    extra = rng.standard_normal((5, 4489)).astype(np.float32) * np.float32(30.0)
    allrows = np.concatenate([rows, extra])
    u8, raw = E.dct_roi_decode(torch.tensor(allrows).cuda(), "per_frame", want_raw=True)
    u8, raw = u8.cpu().numpy(), raw.cpu().numpy()
    ref_raw = np.stack([ov.dct_to_roi(r) for r in allrows])
    assert np.allclose(raw, ref_raw, rtol=1e-6, atol=1e-3 * np.abs(ref_raw).max() * 1e-3)
    ref_u8 = np.stack([ov.roi_to_u8_per_frame(r) for r in ref_raw])
    d = np.abs(u8.astype(np.int32) - ref_u8.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-4, (d.max(), (d != 0).mean())  # rint ties only
    # the reference's own upsampled frames (x264 RGB<->YUV round trip: +-1 grey level)
    gold = gup["sa1_X_first24"].astype(np.int32)
    up = u8[:12][ov.upsample_index(152, 24)].astype(np.int32)
    dd = np.abs(up - gold)
    assert dd.max() <= 1 and (dd == 0).mean() > 0.8


def test_dct_decode_global_matches_oracle():
    rng = np.random.default_rng(1)
    rows = (rng.standard_normal((9, 4489)) * 20.0).astype(np.float32)
    got = E.dct_roi_decode(torch.tensor(rows).cuda(), "global").cpu().numpy()
    ref = ov.roi_to_u8_global(np.stack([ov.dct_to_roi(r) for r in rows]))
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("utt", ["sa1", "sa2", "si494"])
def test_vad_and_ibm_labels_reproduce_reference_files(utt):
    g = golden("golden_frontend_34M.npz")
    x = ofe.peak_normalise(g[utt + "_wav"].astype(np.float32) / 32768.0)
    n = len(x)
    T = E.stft_num_frames(n)
    w = torch.tensor(x, device="cuda")[None]
    vad = E.vad_labels(w, [n], [T], T)[0].cpu().numpy()
    assert np.array_equal(vad.astype(np.uint8), g[utt + "_vad"][0])
    S = E.stft(w, [n], [T], T)
    mask = E.ibm_labels(S, [T])[0].cpu().numpy()
    shape = tuple(g[utt + "_ibm_shape"])
    gold = np.unpackbits(g[utt + "_ibm_bits"])[: shape[0] * shape[1]].reshape(shape)
    assert mask.shape == shape
    assert int((mask.astype(np.uint8) != gold).sum()) == 0


def test_labels_batched_and_ragged():
    rng = np.random.default_rng(3)
    ns = [20000, 31000, 16384]
    B, N = len(ns), max(ns)
    wave = np.zeros((B, N), np.float32)
    for b, n in enumerate(ns):
        wave[b, :n] = rng.standard_normal(n).astype(np.float32) * (0.02 + 0.3 * (np.arange(n) > n // 2))
    Ts = [E.stft_num_frames(n) for n in ns]
    tmax = max(Ts)
    w = torch.tensor(wave).cuda()
    vad = E.vad_labels(w, ns, Ts, tmax).cpu().numpy()
    S = E.stft(w, ns, Ts, tmax)
    ibm = E.ibm_labels(S, Ts).cpu().numpy()
    Sn = S.cpu().numpy()
    for b, n in enumerate(ns):
        ref_vad = ofe.clean_speech_VAD(wave[b, :n])
        assert np.array_equal(vad[b, : Ts[b]], ref_vad[0]) and not vad[b, Ts[b]:].any()
        ref_ibm = ofe.clean_speech_IBM(Sn[b, :, : Ts[b], 0] + 1j * Sn[b, :, : Ts[b], 1])
        assert (ibm[b, :, : Ts[b]] != ref_ibm).mean() < 1e-5 and not ibm[b, :, Ts[b]:].any()


def test_running_statistics_match_the_reference_formula():
    rng = np.random.default_rng(5)
    st = E.RunningStats(513)
    xs, tot = [], 0
    for it in range(3):
        B, T = 4, 50 + 7 * it
        x = (rng.standard_normal((B, T, 513)) * 2.0 - 3.0).astype(np.float32)
        lens = [T, T - 5, 10, T - 1]
        st.update(torch.tensor(x).cuda(), lens)
        for b in range(B):
            xs.append(x[b, : lens[b]])
    mean, std = st.finalize()
    X = np.concatenate(xs).astype(np.float64)            # (n, 513)
    n = X.shape[0]
    m = X.sum(0) / n
    s = np.sqrt((1 / (n - 1)) * ((X ** 2).sum(0) - n * m ** 2))
    assert st.n == n
    assert np.allclose(mean.cpu().numpy(), m, rtol=1e-6, atol=1e-6)
    assert np.allclose(std.cpu().numpy(), s, rtol=1e-6, atol=1e-6)
