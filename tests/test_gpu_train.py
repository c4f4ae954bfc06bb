"""GPU parity: training step of the audio-only model (LSTM + head): loss, gradients (BPTT on the device) and the
fused Adam update against fp32 autograd of the oracle / torch.optim.Adam on the CPU.  bf16 operands: gradients are
held to 3 % relative Frobenius error."""
import copy

import numpy as np
import pytest
import torch

from avvad import engine as E
from avvad import synth
from oracle import models as om
from util import err_stats

pytestmark = pytest.mark.gpu


def _setup(B=5, T=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, 513, generator=g)
    y = (torch.rand(B, T, 1, generator=g) > 0.5).float()
    lens = [24, 17, 24, 9, 1][:B]
    sd = synth.seeded_state_dict(synth.model_spec("audio"), seed=77)
    return x, y, lens, sd


def _oracle_grads(x, y, lens, sd):
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits = om.deepvad_audio_forward(x, lens, p)
    loss = om.batch_loss(logits, y, lens, 1e-8)
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items()}


def test_audio_training_step_gradients_match_autograd():
    from packages.models.Audio_Net import DeepVAD_audio
    x, y, lens, sd = _setup()
    ref_loss, ref_g = _oracle_grads(x, y, lens, sd)
    m = DeepVAD_audio(2, 1024, 1)
    m.load_state_dict(sd)
    m = m.cuda().train()
    logits = m(x.cuda(), torch.tensor(lens).cuda())
    assert logits.requires_grad
    loss, per, dl = E.batch_bce(logits, y.cuda(), lens, 1e-8, want_grad=True)
    assert abs(loss.item() - ref_loss) < 2e-2 * max(1.0, abs(ref_loss))
    logits.backward(dl)
    report = {}
    for k, p in m.named_parameters():
        st = err_stats(p.grad.cpu().numpy(), ref_g[k].numpy())
        report[k] = round(st["rel_fro"], 4)
        assert st["rel_fro"] < 3e-2, (k, st)
    print("grad rel errors:", report)
    # the same through the reference-style Python loss (scripts/train_audio_net.py) + autograd
    m.zero_grad()
    logits = m(x.cuda(), lens)
    from packages.models.utils import binary_cross_entropy
    loss2 = 0.
    for n, pred, tgt in zip(lens, logits, y.cuda()):
        loss2 = loss2 + binary_cross_entropy(pred[:n], tgt[:n], 1e-8)
    loss2.backward()
    st = err_stats(m.lstm_audio.weight_hh_l1.grad.cpu().numpy(), ref_g["lstm_audio.weight_hh_l1"].numpy())
    assert st["rel_fro"] < 3e-2, st


def test_fused_adam_matches_torch_adam():
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(1000, 37, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.9, 0.999))
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(1000, 37, generator=g)
        ref.grad = gr.clone()
        opt.step()
        E.adam_step(p, gr.cuda(), m, v, step, 1e-4, (0.9, 0.999), 1e-8)
    assert np.allclose(p.cpu().numpy(), ref.detach().numpy(), atol=1e-7, rtol=1e-5)


def test_two_training_steps_reduce_the_loss():
    from packages.models.Audio_Net import DeepVAD_audio
    from packages.models._engine import FusedAdam
    x, y, lens, sd = _setup(seed=3)
    m = DeepVAD_audio(2, 1024, 1)
    m.load_state_dict(sd)
    m = m.cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        logits = m(x.cuda(), lens)
        loss, _, dl = E.batch_bce(logits, y.cuda(), lens, 1e-8, want_grad=True)
        logits.backward(dl)
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


def test_trunk_train_mode_bn_matches_oracle():
    """Batch-statistics BatchNorm: features and updated running statistics vs the fp32 oracle."""
    sd = synth.seeded_state_dict(synth.model_spec("video"), seed=12)
    frames = torch.randn(24, 67, 67, generator=torch.Generator().manual_seed(4))
    ref = om.resnet18_trunk(frames, sd, training=True).numpy()
    # oracle running-stat update of the stem BN (momentum 0.1, unbiased variance)
    x0 = torch.nn.functional.conv2d(frames.unsqueeze(1).repeat(1, 3, 1, 1), sd["features.0.weight"], None, stride=2, padding=3)
    rm_ref = 0.9 * sd["features.1.running_mean"] + 0.1 * x0.mean(dim=(0, 2, 3))
    rv_ref = 0.9 * sd["features.1.running_var"] + 0.1 * x0.var(dim=(0, 2, 3), unbiased=True)
    trunk = E.ResNet18Trunk()
    trunk.load_train(sd, "cuda")
    running = []
    for _, bk in E.RESNET_LAYER_KEYS:
        running.append((sd["features." + bk + ".running_mean"].clone().cuda(), sd["features." + bk + ".running_var"].clone().cuda()))
    feat = trunk.forward_train(frames.cuda(), running).cpu().numpy()
    st = err_stats(feat, ref)
    assert st["rel_fro"] < 3e-2, st
    assert np.allclose(running[0][0].cpu().numpy(), rm_ref.numpy(), atol=2e-3)
    assert np.allclose(running[0][1].cpu().numpy(), rv_ref.numpy(), rtol=2e-2, atol=1e-3)


@pytest.mark.parametrize("use_mcb", [False, True])
def test_av_training_step_matches_autograd(use_mcb):
    from packages.models.AV_Net import DeepVAD_AV
    B, T = 3, 10
    lens = [10, 7, 4]
    g = torch.Generator().manual_seed(8)
    a = torch.randn(B, T, 513, generator=g)
    v = torch.randn(B, T, 67, 67, generator=g)
    y = (torch.rand(B, T, 1, generator=g) > 0.5).float()
    sd = synth.seeded_state_dict(synth.model_spec("av", use_mcb=use_mcb), seed=55)
    # oracle: train-mode forward (batch-stat BN in trunk and mcb_bn), trainable = everything but the trunk
    p = {k: (t.clone().requires_grad_(True) if (t.is_floating_point() and not k.startswith("features.")
                                                  and "running" not in k and not k.startswith("mcb.sketch")) else t.clone())
         for k, t in sd.items()}
    logits_ref = om.deepvad_av_forward(a, v, lens, p, use_mcb=use_mcb, eps=1e-8, training=True)
    loss_ref = om.batch_loss(logits_ref, y, lens, 1e-8)
    loss_ref.backward()
    m = DeepVAD_AV(2, 1024, 1, use_mcb=use_mcb, eps=1e-8)
    m.load_state_dict(sd)
    for name, child in m.named_children():  # scripts/train_AV_net.py:241-245
        if name == "features":
            for q in child.parameters():
                q.requires_grad = False
    m = m.cuda().train()
    logits = m(a.cuda(), v.cuda(), torch.tensor(lens).cuda())
    loss, _, dl = E.batch_bce(logits, y.cuda(), lens, 1e-8, want_grad=True)
    assert abs(loss.item() - loss_ref.item()) < 3e-2 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    logits.backward(dl)
    names = ["lstm_merged.weight_ih_l0", "lstm_merged.weight_hh_l0", "lstm_merged.bias_ih_l0", "lstm_merged.weight_ih_l1",
             "lstm_merged.weight_hh_l1", "vad_merged.weight"]
    if use_mcb:
        names += ["mcb_bn.weight", "mcb_bn.bias"]
    report = {}
    params = dict(m.named_parameters())
    for k in names:
        st = err_stats(params[k].grad.cpu().numpy(), p[k].grad.numpy())
        report[k] = round(st["rel_fro"], 4)
    print("AV grad rel errors:", report)
    for k, e in report.items():
        assert e < 6e-2, report
    assert int(m.features[1].num_batches_tracked) == 1


def test_ibm_head_training_gradients_match_autograd():
    """labels = 'ibm_labels' (y_dim = 513, scripts/train_AV_net.py:65-66): the head backward is three GEMMs instead of
    the rank-1 VAD path; loss and every gradient against fp32 autograd of the oracle."""
    from packages.models.Audio_Net import DeepVAD_audio
    B, T = 4, 20
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, 513, generator=g)
    y = (torch.rand(B, T, 513, generator=g) > 0.5).float()
    lens = [20, 13, 20, 6]
    sd = synth.seeded_state_dict(synth.model_spec("audio", y_dim=513), seed=31)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_logits = om.deepvad_audio_forward(x, lens, p)
    ref_loss = om.batch_loss(ref_logits, y, lens, 1e-8)
    ref_loss.backward()
    m = DeepVAD_audio(2, 1024, 513)
    m.load_state_dict(sd)
    m = m.cuda().train()
    logits = m(x.cuda(), lens)
    assert logits.shape == (B, T, 513) and logits.requires_grad
    loss, per, dl = E.batch_bce(logits, y.cuda(), lens, 1e-8, want_grad=True)
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    logits.backward(dl)
    report = {}
    for k, q in m.named_parameters():
        st = err_stats(q.grad.cpu().numpy(), p[k].grad.numpy())
        report[k] = round(st["rel_fro"], 4)
        assert st["rel_fro"] < 3e-2, (k, st)
    print("ibm-head grad rel errors:", report)


def test_video_training_step_with_frozen_trunk_matches_autograd():
    """DeepVAD_video in train() with `features` frozen (train_AV_net.py:241-245 applied to the video-only model):
    batch-statistics trunk + device BPTT vs fp32 autograd of the oracle (the trainable trunk is covered by
    tests/test_gpu_trunk_backward.py)."""
    from packages.models.Video_Net import DeepVAD_video
    B, T = 3, 10
    lens = [10, 7, 4]
    g = torch.Generator().manual_seed(18)
    v = torch.randn(B, T, 67, 67, generator=g)
    y = (torch.rand(B, T, 1, generator=g) > 0.5).float()
    sd = synth.seeded_state_dict(synth.model_spec("video"), seed=56)
    p = {k: (t.clone().requires_grad_(True) if (t.is_floating_point() and not k.startswith("features.")) else t.clone())
         for k, t in sd.items()}
    logits_ref = om.deepvad_video_forward(v, lens, p, training=True)
    loss_ref = om.batch_loss(logits_ref, y, lens, 1e-8)
    loss_ref.backward()
    m = DeepVAD_video(2, 1024, 1)
    m.load_state_dict(sd)
    m = m.cuda().train()
    for q in m.features.parameters():
        q.requires_grad = False
    logits = m(v.cuda(), torch.tensor(lens).cuda())
    loss, _, dl = E.batch_bce(logits, y.cuda(), lens, 1e-8, want_grad=True)
    assert abs(loss.item() - loss_ref.item()) < 3e-2 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    logits.backward(dl)
    params = dict(m.named_parameters())
    report = {}
    for k in ("lstm_video.weight_ih_l0", "lstm_video.weight_hh_l0", "lstm_video.bias_ih_l0", "lstm_video.weight_ih_l1",
              "lstm_video.weight_hh_l1", "vad_video.weight", "vad_video.bias"):
        report[k] = round(err_stats(params[k].grad.cpu().numpy(), p[k].grad.numpy())["rel_fro"], 4)
    print("video grad rel errors:", report)
    assert all(e < 6e-2 for e in report.values()), report
    assert int(m.features[1].num_batches_tracked) == 1
    m.eval()
    with torch.no_grad():
        out = m(v.cuda(), lens)
    assert out.shape == (B, T, 1) and not out.requires_grad
