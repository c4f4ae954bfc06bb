"""GPU parity: fused masked BCE (+ gradient) and per-utterance f1 metrics vs the oracle / reference golden values."""
import numpy as np
import pytest
import torch

from avvad import engine as E
from oracle import models as om
from util import golden

pytestmark = pytest.mark.gpu


def _case(B=5, T=40, seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, T, 1, generator=g) * 3
    target = (torch.rand(B, T, 1, generator=g) > 0.4).float()
    lens = [40, 33, 1, 17, 25][:B]
    return logits, target, lens


def test_batch_bce_matches_oracle_and_autograd():
    logits, target, lens = _case()
    lr = logits.clone().requires_grad_(True)
    ref = om.batch_loss(lr, target, lens, 1e-8)
    ref.backward()
    loss, per, grad = E.batch_bce(logits.cuda(), target.cuda(), lens, 1e-8, want_grad=True)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert np.allclose(grad.cpu().numpy(), lr.grad.numpy(), atol=1e-6, rtol=1e-4)
    for b, n in enumerate(lens):
        assert abs(per[b].item() - om.binary_cross_entropy(logits[b, :n], target[b, :n], 1e-8).item()) < 1e-5
        assert torch.all(grad[b, n:] == 0)


def test_bce_golden_from_reference():
    g = golden("ref_models.npz")
    r, t = torch.tensor(g["bce_r"])[None], torch.tensor(g["bce_t"])[None]  # (1,37,1)
    loss, _, _ = E.batch_bce(r.cuda(), t.cuda(), [37], 1e-8)
    assert abs(loss.item() - float(g["bce_out"])) < 1e-5


def test_batch_f1_matches_oracle_and_reference_golden():
    logits, target, lens = _case(seed=3)
    met, dec = E.batch_f1(logits.cuda(), target.cuda(), lens, 1e-8)
    for b, n in enumerate(lens):
        hard = (torch.sigmoid(logits[b, :n, 0]) > 0.5).int()
        ref = [v.item() for v in om.f1_loss(hard, target[b, :n, 0].long(), 1e-8)]
        assert np.allclose(met[b].cpu().numpy(), ref, atol=1e-6), (b, met[b], ref)
        assert torch.equal(dec[b, :n].cpu(), hard)
    g = golden("ref_models.npz")
    r, t = torch.tensor(g["bce_r"])[None], torch.tensor(g["bce_t"])[None]
    met, _ = E.batch_f1(r.cuda(), t.cuda(), [37], 1e-8)
    assert np.allclose(met[0].cpu().numpy(), g["f1_out"], atol=1e-6)
