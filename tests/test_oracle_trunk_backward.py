"""CPU: pin the from-tape trunk-backward oracle (oracle/trunk_backward.py) against plain PyTorch autograd of the
training-mode ResNet-18 trunk (the semantics of packages/models/Video_Net.py:60-99 under scripts/train_video_net.py).
The "tape" here is built from an fp32 forward, so the two must agree to rounding."""
import torch
import torch.nn.functional as F

from avvad import synth
from oracle.trunk_backward import BLOCKS, LAYERS, trunk_backward_from_tape


def _forward_collect(frames, p):
    def nhwc(t):
        return t.detach().permute(0, 2, 3, 1).contiguous()
    saved = {"raw": [None] * 20, "y1": [None] * 8, "out": [None] * 8}
    stats = torch.zeros(20, 1024)

    def bn(x, l):
        ck, bk, ci, co, k, s, pd = LAYERS[l]
        saved["raw"][l] = nhwc(x)
        stats[l, :co] = x.mean((0, 2, 3)).detach()
        stats[l, co:2 * co] = (1 / torch.sqrt(x.var((0, 2, 3), unbiased=False) + 1e-5)).detach()
        return F.batch_norm(x, None, None, p["features." + bk + ".weight"], p["features." + bk + ".bias"], True, 0.1, 1e-5)

    x = frames.unsqueeze(1).repeat(1, 3, 1, 1)
    x = F.relu(bn(F.conv2d(x, p["features.0.weight"], None, 2, 3), 0))
    saved["act0"] = nhwc(x)
    x = F.max_pool2d(x, 3, 2, 1)
    saved["pool"] = nhwc(x)
    for bk, (la, lb, lds) in enumerate(BLOCKS):
        s = LAYERS[la][5]
        y = F.relu(bn(F.conv2d(x, p["features." + LAYERS[la][0] + ".weight"], None, s, 1), la))
        saved["y1"][bk] = nhwc(y)
        z = bn(F.conv2d(y, p["features." + LAYERS[lb][0] + ".weight"], None, 1, 1), lb)
        idt = x if lds < 0 else bn(F.conv2d(x, p["features." + LAYERS[lds][0] + ".weight"], None, s, 0), lds)
        x = F.relu(z + idt)
        saved["out"][bk] = nhwc(x)
    saved["stats"] = stats
    return F.adaptive_avg_pool2d(x, 1).flatten(1), saved


def test_from_tape_backward_equals_autograd():
    torch.manual_seed(0)
    n = 6
    frames = torch.randn(n, 67, 67)
    sd = synth.seeded_state_dict(synth.model_spec("video"), 61, "strong")
    p = {k: (t.clone().requires_grad_(True) if t.is_floating_point() and "running" not in k else t.clone())
         for k, t in sd.items()}
    feat, saved = _forward_collect(frames, p)
    dfeat = torch.randn(n, 512)
    feat.backward(dfeat)
    g = trunk_backward_from_tape(frames, saved, sd, dfeat, weights_bf16=False)
    assert len(g) == 60
    for k, v in g.items():
        ref = p[k].grad
        assert ((v - ref).norm() / ref.norm()).item() < 1e-4, k
