"""GPU parity, end to end: raw waveforms + 30 fps ROI frames through AVVADPipeline vs the CPU port of
the reference path (torch.stft front end, index-map upsampling, torchvision ResNet-18, torch.fft MCB,
nn.LSTM, Linear, sigmoid, threshold)."""
import numpy as np
import pytest
import torch

from avvad import synth
from avvad.pipeline import AVVADPipeline
from oracle.reference_port import RefDeepVADAV, cpu_av_step

pytestmark = pytest.mark.gpu


def _inputs(B, seed=0):
    ns = [24000 + 1777 * i for i in range(B)]           # ragged: 1.5 s .. 2+ s
    nf = [45 + 3 * i for i in range(B)]                 # 30 fps frames
    waves = [synth.synth_wave(n, seed + i) for i, n in enumerate(ns)]
    vids = [synth.synth_video_u8(f, seed + i) for i, f in enumerate(nf)]
    return ns, nf, waves, vids


@pytest.mark.parametrize("use_mcb,piece", [(False, 32), (True, 32), (True, 3), (False, 1)])
def test_pipeline_matches_reference_port(use_mcb, piece):
    B = 4
    ns, nf, waves, vids = _inputs(B)
    mean, std = synth.synth_audio_stats(0)
    sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=use_mcb), seed=41)
    lens = AVVADPipeline.frame_counts(ns, nf)
    tmax = max(lens)
    if use_mcb:
        sd["mcb_bn.running_mean"] = torch.zeros(1024)
        sd["mcb_bn.running_var"] = torch.full((1024,), 1.0 / (B * tmax * 1024.0))
    ref_model = RefDeepVADAV(2, 1024, 1, use_mcb=use_mcb).load_reference_state_dict(sd).eval()
    rpost, rdec, rlens = cpu_av_step(ref_model, waves, vids, mean, std, synth.VIDEO_MEAN, synth.VIDEO_STD)
    assert rlens == lens

    pipe = AVVADPipeline(sd, mean, std, synth.VIDEO_MEAN, synth.VIDEO_STD, use_mcb=use_mcb)
    pipe.piece = piece  # utterances per upload/trunk piece: 32 = one piece, 3 and 1 exercise the overlapped path
    wave = torch.zeros(B, max(ns))
    vid = torch.zeros(B, max(nf), 67, 67, dtype=torch.uint8)
    for i in range(B):
        wave[i, : ns[i]] = torch.from_numpy(waves[i])
        vid[i, : nf[i]] = torch.from_numpy(vids[i])
    post, dec = pipe.infer_host(wave.pin_memory(), ns, vid.pin_memory(), nf)
    post, dec = post[..., 0].numpy(), dec[..., 0].numpy()
    for b in range(B):
        d = np.abs(post[b, : lens[b]] - rpost[b, : lens[b]].numpy())
        assert d.max() < 1e-2, (b, d.max())
        sure = np.abs(rpost[b, : lens[b]].numpy() - 0.5) > 1e-2
        assert np.array_equal(dec[b, : lens[b]][sure], rdec[b, : lens[b]].numpy()[sure])
    # padded steps carry sigmoid(bias) exactly like the reference (zeros through the Linear)
    bias = float(sd["vad_merged.bias"][0])
    for b in range(B):
        if lens[b] < tmax:
            assert np.allclose(post[b, lens[b]:], 1.0 / (1.0 + np.exp(-bias)), atol=1e-6)


@pytest.mark.parametrize("use_mcb", [False, True])
def test_dedup_video_is_bit_identical(use_mcb):
    """Trunk on the 30 fps source frames + feature gather == trunk on every upsampled frame (ragged lengths, padded
    frames included): the optional dedup path must not change a single bit of the posteriors."""
    B = 5
    ns, nf, waves, vids = _inputs(B, seed=7)
    nf[2] = 20  # a short video: many collate-padded frames
    vids[2] = vids[2][:20]
    mean, std = synth.synth_audio_stats(0)
    sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=use_mcb), seed=5)
    wave = torch.zeros(B, max(ns))
    vid = torch.zeros(B, max(nf), 67, 67, dtype=torch.uint8)
    for i in range(B):
        wave[i, : ns[i]] = torch.from_numpy(waves[i])
        vid[i, : nf[i]] = torch.from_numpy(vids[i])
    wave, vid = wave.cuda(), vid.cuda()
    lens = [min(a, b) for a, b in zip(AVVADPipeline.frame_counts(ns, nf), [AVVADPipeline.frame_counts(ns, nf)[i] for i in range(B)])]
    tmax = max(lens) + 3  # a few all-padding columns as well
    outs = []
    for dedup in (False, True):
        pipe = AVVADPipeline(sd, mean, std, synth.VIDEO_MEAN, synth.VIDEO_STD, use_mcb=use_mcb)
        pipe.dedup_video = dedup
        pipe.piece = 2
        logits, post, dec = pipe.infer_device(wave, ns, vid, nf, lengths=lens, t_max=tmax)
        outs.append((logits.clone(), post.clone(), dec.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
