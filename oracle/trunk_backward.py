"""Oracle: back-propagation through the ResNet-18 trunk GIVEN the forward's saved tensors (fp32 PyTorch, CPU).  Test
infrastructure only.

Why "given the saved tensors": the training-mode trunk at random initialisation is chaotic with respect to rounding --
the fp32 autograd gradient of a forward whose tensors are rounded to bf16 differs by 14-25 % (relative Frobenius norm,
every conv layer) between fp32 and fp64 ARITHMETIC of that same model (oracle.models.bf16_ste; 40-frame batch), because
a last-bit change flips ReLU masks and moves BatchNorm batch statistics, and the flips compound through 17 layers.  An
end-to-end gradient comparison therefore cannot distinguish a correct device backward from a wrong one.  This module
instead evaluates the textbook backward formulas, layer by layer, on the activations the device's forward actually
produced (its tape: raw convolution outputs, post-activation tensors, batch mean / inverse std), chaining its own fp32
gradients from the top -- so the only admissible difference to the device result is the bf16 storage of the gradient
activations.

Follows autograd of torchvision's BasicBlock / BatchNorm2d(train) / MaxPool2d / AdaptiveAvgPool2d as used by
packages/models/Video_Net.py:60-99 (scripts/train_video_net.py:145-173 trains the trunk)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

# (conv key, bn key, cin, cout, k, stride, pad) in libavvad's conv-layer order (include/avvad.h)
LAYERS = [("0", "1", 3, 64, 7, 2, 3)]
for _st, _ci, _co in ((4, 64, 64), (5, 64, 128), (6, 128, 256), (7, 256, 512)):
    _s = 1 if _st == 4 else 2
    LAYERS += [(f"{_st}.0.conv1", f"{_st}.0.bn1", _ci, _co, 3, _s, 1), (f"{_st}.0.conv2", f"{_st}.0.bn2", _co, _co, 3, 1, 1)]
    if _st != 4:
        LAYERS += [(f"{_st}.0.downsample.0", f"{_st}.0.downsample.1", _ci, _co, 1, 2, 0)]
    LAYERS += [(f"{_st}.1.conv1", f"{_st}.1.bn1", _co, _co, 3, 1, 1), (f"{_st}.1.conv2", f"{_st}.1.bn2", _co, _co, 3, 1, 1)]
BLOCKS = [(1, 2, -1), (3, 4, -1), (5, 6, 7), (8, 9, -1), (10, 11, 12), (13, 14, -1), (15, 16, 17), (18, 19, -1)]


def _nchw(t):  # tape tensors are NHWC bf16
    return t.float().permute(0, 3, 1, 2).contiguous()


def _bn_backward(raw, g, mean, invstd, gamma):
    """dRaw, dgamma, dbeta of y = gamma * (raw - mean) * invstd + beta with batch statistics (mean, invstd)."""
    xh = (raw - mean[None, :, None, None]) * invstd[None, :, None, None]
    m = raw.numel() / raw.shape[1]
    s1 = g.sum((0, 2, 3))
    s2 = (g * xh).sum((0, 2, 3))
    d = (gamma * invstd)[None, :, None, None] * (g - s1[None, :, None, None] / m - xh * s2[None, :, None, None] / m)
    return d, s2, s1


def trunk_backward_from_tape(frames: torch.Tensor, saved: Dict, sd: Dict[str, torch.Tensor], dfeat: torch.Tensor,
                             prefix="features.", weights_bf16=True) -> Dict[str, torch.Tensor]:
    """frames (n,67,67) f32; saved = ResNet18Trunk.tape_tensors(...) moved to the CPU; dfeat (n,512).
    Returns {state_dict key: gradient} for the 20 conv weights and 40 BatchNorm affine parameters."""
    n = frames.shape[0]
    raw = [_nchw(t) for t in saved["raw"]]
    y1 = [_nchw(t) for t in saved["y1"]]
    out = [_nchw(t) for t in saved["out"]]
    act0, pool = _nchw(saved["act0"]), _nchw(saved["pool"])
    stats = saved["stats"].float()
    grads: Dict[str, torch.Tensor] = {}

    def W(l):  # the device convolves with bf16 copies of the weights (conv1 stays fp32)
        w = sd[prefix + LAYERS[l][0] + ".weight"].float()
        return w.to(torch.bfloat16).float() if (weights_bf16 and l > 0) else w

    def bn(l, g):
        ck, bk, ci, co, k, s, p = LAYERS[l]
        d, dg, db = _bn_backward(raw[l], g, stats[l, :co], stats[l, co:2 * co], sd[prefix + bk + ".weight"].float())
        grads[prefix + bk + ".weight"], grads[prefix + bk + ".bias"] = dg, db
        return d

    def conv_grads(l, x, d, need_input=True):
        ck, bk, ci, co, k, s, p = LAYERS[l]
        grads[prefix + ck + ".weight"] = torch.nn.grad.conv2d_weight(x, (co, ci, k, k), d, stride=s, padding=p)
        return torch.nn.grad.conv2d_input(x.shape, W(l), d, stride=s, padding=p) if need_input else None

    g = (dfeat.float() / 9.0)[:, :, None, None].expand(-1, -1, 3, 3)
    for bk in range(7, -1, -1):
        la, lb, lds = BLOCKS[bk]
        x = pool if bk == 0 else out[bk - 1]
        gm = g * (out[bk] > 0)
        d_b = bn(lb, gm)
        g_y1 = conv_grads(lb, y1[bk], d_b)
        d_a = bn(la, g_y1 * (y1[bk] > 0))
        g_x = conv_grads(la, x, d_a)
        if lds >= 0:
            d_d = bn(lds, gm)
            g_x = g_x + conv_grads(lds, x, d_d)
        else:
            g_x = g_x + gm
        g = g_x
    # stem: max-pool (first maximum) -> ReLU mask -> BatchNorm -> conv1 on the tripled single-channel frame
    a0 = act0.clone().requires_grad_(True)
    F.max_pool2d(a0, 3, 2, 1).backward(g)
    d0 = bn(0, a0.grad * (act0 > 0))
    x0 = frames.float().unsqueeze(1).repeat(1, 3, 1, 1)
    grads[prefix + "0.weight"] = torch.nn.grad.conv2d_weight(x0, (64, 3, 7, 7), d0, stride=2, padding=3)
    return grads
