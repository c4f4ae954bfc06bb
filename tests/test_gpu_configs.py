"""BASELINE.json configs[0] and configs[1] on the reference's own data (tests/golden holds the utterance test/34M/sa1 of
data/subset: noisy waveform, shipped standardisation statistics, the first upsampled mouth-ROI frames):
  0. audio-only VAD forward (scripts/evaluate_audio_net.py path): waveform -> STFT/log-power/standardise -> LSTM -> head
  1. video-only VAD forward on upsampled ROI frames (scripts/evaluate_video_net.py path)
GPU path vs the CPU oracle with the same seeded weights (the reference ships no checkpoints; "strong" family so that the
logits span several units, head bias placed by synth.decision_bias): logits within 2e-2 relative, frame posteriors within
1e-2, decisions identical on >= 99.9 % of ALL frames, accuracy (packages/models/utils.f1_loss) within 0.001."""
import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from oracle import models as om
from avvad import engine as E
from avvad import synth
from util import golden, check_logits, sigmoid as _sig

pytestmark = pytest.mark.gpu


def _place_bias(module, head, forward_oracle, lens):
    """Two-pass head-bias placement from the ORACLE's logits (logits are affine in the bias)."""
    sd = {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}
    ref0 = forward_oracle(sd)
    key = [k for k in sd if k.startswith("vad_") and k.endswith(".bias")][0]
    nb = synth.decision_bias(ref0, lens, sd[key].numpy())
    ref = ref0 - sd[key].numpy() + nb.numpy()
    with torch.no_grad():
        head.bias.copy_(nb)
    sd[key] = nb
    return sd, ref


def test_config0_audio_only_forward_on_subset_utterance():
    from packages.models.Audio_Net import DeepVAD_audio
    from packages.models.utils import f1_loss
    g = golden("golden_frontend_34M.npz")
    wav = g["sa1_noisy_wav"].astype(np.float32) / 32768.0
    mean, std = g["audio_mean"][:, 0], g["audio_std"][:, 0]
    n = len(wav)
    T = E.stft_num_frames(n)
    assert T == 317
    # oracle: the reference's host path in float32 (peak normalise, STFT, log-power, standardise)
    feat_ref = ofe.frontend_features(wav, mean, std, dtype=np.float32)          # (T, 513)
    x_ref = torch.tensor(np.ascontiguousarray(feat_ref))[None]                  # (1, T, 513)
    m = synth.fill_module_(DeepVAD_audio(2, 1024, 1), seed=101, family="strong")
    sd, ref = _place_bias(m, m.vad_audio, lambda q: om.deepvad_audio_forward(x_ref, [T], q).numpy(), [T])
    m = m.cuda().eval()
    # device path: raw waveform in
    w = torch.tensor(wav, device="cuda")[None]
    feat = E.frontend_logpower(w, [n], [T], T, torch.tensor(mean).cuda(), torch.tensor(std).cuda(), 1e-8, True)
    with torch.no_grad():
        logits, post, dec = m(feat, [T], return_posteriors=True)
    check_logits(logits.cpu().numpy(), ref, [T], "config 0 (audio-only, test/34M/sa1)")
    post, dec = post[0, :, 0].cpu().numpy(), dec[0, :, 0].cpu().numpy()
    assert np.array_equal(dec, (post > 0.5).astype(np.int32))
    # accuracy against the shipped VAD labels of the clean utterance, both paths (within 0.001)
    lab = torch.tensor(g["sa1_vad"][0].astype(np.int64))
    acc_dev = f1_loss(torch.tensor(dec.astype(np.int64)), lab)[0].item()
    acc_ref = f1_loss(torch.tensor((ref[0, :, 0] > 0).astype(np.int64)), lab)[0].item()
    assert abs(acc_dev - acc_ref) <= 1e-3


def test_config1_video_only_forward_on_upsampled_subset_frames():
    from packages.models.Video_Net import DeepVAD_video
    g = golden("golden_upsample.npz")
    gf = golden("golden_frontend_34M.npz")
    frames_u8 = g["sa1_X_first24"]                                               # (24,67,67) from the shipped *_upsampled.h5
    vm, vs = float(gf["video_mean"][0, 0]), float(gf["video_std"][0, 0])
    x = ((frames_u8.astype(np.float32) - np.float32(vm)) / (np.float32(vs) + np.float32(1e-8)))[None]  # (1,24,67,67)
    m = synth.fill_module_(DeepVAD_video(2, 1024, 1), seed=102, family="strong")
    sd, ref = _place_bias(m, m.vad_video, lambda q: om.deepvad_video_forward(torch.tensor(x), [24], q).numpy(), [24])
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(torch.tensor(x).cuda(), [24]).cpu().numpy()
    check_logits(out, ref, [24], "config 1 (video-only, first 24 shipped frames of test/34M/sa1)")
    # the fused u8 path (gather + standardise inside the stem) on the same frames: identity index map
    trunk = E.ResNet18Trunk()
    trunk.load(sd, "cuda")
    fa = trunk.forward(torch.tensor(x[0]).cuda())
    fb = trunk.forward_u8(torch.tensor(frames_u8)[None].cuda(), [24], [24], 24, vm, vs, 1e-8, True, num=1, den=1)
    assert torch.equal(fa, fb)
