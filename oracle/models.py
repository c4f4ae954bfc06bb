"""Oracle: model forward / loss / metrics in fp32 PyTorch on CPU.  Test infrastructure only.

Functional restatements driven by a ``state_dict`` (same keys as the reference modules):
  * ResNet-18 trunk (V1-V2) ......... packages/models/AV_Net.py:25-30,78-94 (torchvision resnet18
                                      children[:-1]; not vendored in the reference -- restated from
                                      its published architecture: conv7x7/2+BN+ReLU, maxpool3x3/2,
                                      4 stages x 2 BasicBlocks, global avg-pool)
  * count sketch / MCB (F2) ......... packages/models/compact_bilinear_pooling.py:7-27,140-173
  * signed sqrt, L2, BN1d (F3) ...... packages/models/AV_Net.py:111-121
  * concat fusion (F1) .............. packages/models/AV_Net.py:124
  * packed 2-layer LSTM (R1) ........ packages/models/AV_Net.py:127-137 (nn.LSTM semantics: gate
                                      order i,f,g,o; zero initial state; zero output past len_b)
  * last-step gather (R2) ........... packages/models/utils.py:36-55
  * Linear head (H1) ................ packages/models/AV_Net.py:140
  * sigmoid / threshold (H2) ........ scripts/evaluate_AV_net.py:239-240
  * BCE (L1) ........................ packages/models/utils.py:108-113 + scripts/train_AV_net.py:298-301
  * f1_loss (L2) .................... packages/models/utils.py:164-203
  * WaveNet encoder (W1) ............ packages/models/wavenet_autoencoder.py:74-93
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ---------------------------------------------------------------------------------------------
# ResNet-18 trunk
# ---------------------------------------------------------------------------------------------
def _bn2d(x, sd: SD, p: str, training=False, eps=1e-5):
    if training:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, eps)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.1, eps)


def bf16_ste(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 in the forward, identity in the backward (straight-through).  `quant=bf16_ste` makes the trunk
    below the fp32-autograd model of a network whose FORWARD stores its activations and weights in bf16 -- the right
    yard-stick for the device's trunk backward: at random initialisation the fp32 gradient of the bf16-rounded forward
    already differs from the gradient of the pure fp32 forward by 18 % (last block) to 36 % (stem) in relative Frobenius
    norm (a 0.5 % activation error flips ReLU masks and moves BatchNorm's batch statistics; measured with this function
    on a 40-frame batch), so a comparison against pure fp32 cannot tell a correct backward from a wrong one."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def _basic_block(x, sd: SD, p: str, stride: int, training=False, quant=None):
    q = quant or (lambda t: t)
    out = q(F.conv2d(x, q(sd[p + ".conv1.weight"]), None, stride=stride, padding=1))
    out = q(F.relu(_bn2d(out, sd, p + ".bn1", training)))
    out = q(F.conv2d(out, q(sd[p + ".conv2.weight"]), None, stride=1, padding=1))
    out = _bn2d(out, sd, p + ".bn2", training)
    if (p + ".downsample.0.weight") in sd:
        idt = q(F.conv2d(x, q(sd[p + ".downsample.0.weight"]), None, stride=stride))
        idt = q(_bn2d(idt, sd, p + ".downsample.1", training))
    else:
        idt = x
    return q(F.relu(out + idt))


def resnet18_trunk(frames: torch.Tensor, sd: SD, prefix="features.", training=False,
                   return_intermediates=False, quant=None):
    """frames (M,H,W) single-channel -> (M,512).  The reference triples the channel
    (AV_Net.py:82) and runs torchvision's resnet18 without its fc layer.  `quant` (None = the reference's fp32
    arithmetic) is applied where the device path stores a tensor in bf16: conv outputs, post-activation tensors, conv
    weights (not conv1's, which the training stem keeps in fp32)."""
    q = quant or (lambda t: t)
    x = frames.unsqueeze(1).repeat(1, 3, 1, 1)
    inter = {}
    x = q(F.conv2d(x, sd[prefix + "0.weight"], None, stride=2, padding=3))
    x = q(F.relu(_bn2d(x, sd, prefix + "1", training)))
    inter["conv1"] = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    inter["pool"] = x
    for li, stage in enumerate((4, 5, 6, 7)):
        for blk in (0, 1):
            stride = 2 if (li > 0 and blk == 0) else 1
            x = _basic_block(x, sd, f"{prefix}{stage}.{blk}", stride, training, quant)
            inter[f"l{li + 1}b{blk}"] = x
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    if return_intermediates:
        return x, inter
    return x


# ---------------------------------------------------------------------------------------------
# Fusion
# ---------------------------------------------------------------------------------------------
def count_sketch(x: torch.Tensor, h: torch.Tensor, s: torch.Tensor, out_size: int) -> torch.Tensor:
    """out[..., h_i] += s_i * x[..., i]  (compact_bilinear_pooling.py:7-27)."""
    xs = x * s.view((1,) * (x.dim() - 1) + (-1,))
    out = x.new_zeros(x.shape[:-1] + (out_size,))
    return out.scatter_add_(-1, h.view((1,) * (x.dim() - 1) + (-1,)).expand_as(x), xs)


def mcb(x: torch.Tensor, y: torch.Tensor, sd: SD, prefix="mcb.", out_size=1024) -> torch.Tensor:
    """irfft(rfft(sketch1(x)) * rfft(sketch2(y)))  (compact_bilinear_pooling.py:140-173; the
    legacy torch.rfft/irfft there used the un-normalised forward and 1/N inverse)."""
    px = count_sketch(x, sd[prefix + "sketch1.h"], sd[prefix + "sketch1.s"], out_size)
    py = count_sketch(y, sd[prefix + "sketch2.h"], sd[prefix + "sketch2.s"], out_size)
    return torch.fft.irfft(torch.fft.rfft(px) * torch.fft.rfft(py), n=out_size)


def mcb_fusion(audio, video_feat, sd: SD, eps=1e-8, training=False) -> torch.Tensor:
    """AV_Net.py:111-121: MCB -> signed sqrt -> whole-tensor L2 (detached) -> BatchNorm1d over the
    1024 channels (the two permutes only move the channel axis to dim 1 and back)."""
    y = mcb(audio, video_feat, sd)
    y = torch.sign(y) * torch.sqrt(torch.abs(y) + eps)
    y = y / torch.norm(y, p=2).detach()
    y = y.permute(1, 2, 0).contiguous()
    if training:
        y = F.batch_norm(y, None, None, sd["mcb_bn.weight"], sd["mcb_bn.bias"], True, 0.1, eps)
    else:
        y = F.batch_norm(y, sd["mcb_bn.running_mean"], sd["mcb_bn.running_var"], sd["mcb_bn.weight"],
                         sd["mcb_bn.bias"], False, 0.1, eps)
    return y.permute(2, 0, 1).contiguous()


# ---------------------------------------------------------------------------------------------
# LSTM + head
# ---------------------------------------------------------------------------------------------
def lstm_packed(x: torch.Tensor, lengths: Sequence[int], sd: SD, prefix: str, layers=2,
                return_state=False):
    """Unidirectional multi-layer LSTM over (B,T,I) with per-sequence lengths; outputs are exactly
    zero for t >= len_b (pad_packed_sequence, AV_Net.py:137)."""
    B, T, _ = x.shape
    lengths = [int(v) for v in lengths]
    inp = x
    last_h = None
    for l in range(layers):
        w_ih, w_hh = sd[f"{prefix}.weight_ih_l{l}"], sd[f"{prefix}.weight_hh_l{l}"]
        b = sd[f"{prefix}.bias_ih_l{l}"] + sd[f"{prefix}.bias_hh_l{l}"]
        H = w_hh.shape[1]
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        outs = []
        mask_len = torch.tensor(lengths)
        for t in range(T):
            g = inp[:, t] @ w_ih.t() + h @ w_hh.t() + b
            i, f, gg, o = g.chunk(4, dim=1)
            c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h_new = torch.sigmoid(o) * torch.tanh(c_new)
            m = (mask_len > t).to(x.dtype).unsqueeze(1)
            c = m * c_new + (1 - m) * c
            h = m * h_new + (1 - m) * h
            outs.append(m * h_new)
        inp = torch.stack(outs, dim=1)
        last_h = h
    if return_state:
        return inp, last_h
    return inp


def head(h: torch.Tensor, sd: SD, prefix: str) -> torch.Tensor:
    return F.linear(h, sd[prefix + ".weight"], sd[prefix + ".bias"])


def deepvad_audio_forward(x, lengths, sd: SD, layers=2) -> torch.Tensor:
    """packages/models/Audio_Net.py:43-60."""
    return head(lstm_packed(x, lengths, sd, "lstm_audio", layers), sd, "vad_audio")


def deepvad_video_forward(video, lengths, sd: SD, layers=2, return_last=False, training=False, quant=None) -> torch.Tensor:
    """packages/models/Video_Net.py:58-117."""
    B, T, H, W = video.shape
    feat = resnet18_trunk(video.reshape(B * T, H, W), sd, training=training, quant=quant).view(B, T, -1)
    if return_last:
        _, last = lstm_packed(feat, lengths, sd, "lstm_video", layers, return_state=True)
        return head(last, sd, "vad_video")
    return head(lstm_packed(feat, lengths, sd, "lstm_video", layers), sd, "vad_video")


def deepvad_av_forward(audio, video, lengths, sd: SD, use_mcb=False, eps=1e-8, layers=2,
                       training=False, quant=None) -> torch.Tensor:
    """packages/models/AV_Net.py:72-141."""
    B, T, H, W = video.shape
    feat = resnet18_trunk(video.reshape(B * T, H, W), sd, training=training, quant=quant).view(B, T, -1)
    if use_mcb:
        y = mcb_fusion(audio, feat, sd, eps, training)
    else:
        y = torch.cat([audio, feat], dim=2)
    return head(lstm_packed(y, lengths, sd, "lstm_merged", layers), sd, "vad_merged")


def posteriors_and_decisions(logits: torch.Tensor):
    """scripts/evaluate_AV_net.py:239-240."""
    soft = torch.sigmoid(logits)
    return soft, (soft > 0.5).int()


# ---------------------------------------------------------------------------------------------
# loss / metrics
# ---------------------------------------------------------------------------------------------
def binary_cross_entropy(r, x, eps):
    """packages/models/utils.py:113."""
    return -torch.mean(x * torch.log(torch.sigmoid(r) + eps) + (1 - x) * torch.log(1 - torch.sigmoid(r) + eps))


def batch_loss(logits, target, lengths, eps=1e-8):
    """scripts/train_AV_net.py:298-301: SUM over utterances of the per-utterance mean BCE."""
    loss = logits.new_zeros(())
    for length, pred, tgt in zip(lengths, logits, target):
        loss = loss + binary_cross_entropy(pred[:int(length)], tgt[:int(length)].long().to(pred.dtype), eps)
    return loss


def f1_loss(y_hat_hard, y, epsilon=1e-8):
    """packages/models/utils.py:164-203 -> (accuracy, precision, recall, f1)."""
    y_pred, y_true = y_hat_hard, y
    tp = (y_true * y_pred).sum().to(torch.float32)
    tn = ((1 - y_true) * (1 - y_pred)).sum().to(torch.float32)
    fp = ((1 - y_true) * y_pred).sum().to(torch.float32)
    fn = (y_true * (1 - y_pred)).sum().to(torch.float32)
    accuracy = (tp + tn) / (tp + tn + fp + fn + epsilon)
    precision = tp / (tp + fp + epsilon)
    recall = tp / (tp + fn + epsilon)
    f1 = 2 * (precision * recall) / (precision + recall + epsilon)
    return accuracy, precision, recall, f1


# ---------------------------------------------------------------------------------------------
# WaveNet encoder (dead code in the reference, on the path by north_star decree)
# ---------------------------------------------------------------------------------------------
def wavenet_encode(x, sd: SD, dilations: Sequence[int], pool: int) -> torch.Tensor:
    """packages/models/wavenet_autoencoder.py:74-93.  x (B,q,N) -> (B,bottleneck,pool)."""
    def b(name):
        return sd.get(name + ".bias")

    s = F.conv1d(x, sd["en_causal_layer.weight"], b("en_causal_layer"))
    for i, d in enumerate(dilations):
        cur = s
        s = F.relu(s)
        s = F.conv1d(s, sd[f"en_dilation_layer_stack.{i}.weight"], b(f"en_dilation_layer_stack.{i}"), dilation=d)
        s = F.relu(s)
        s = F.conv1d(s, sd[f"en_dense_layer_stack.{i}.weight"], b(f"en_dense_layer_stack.{i}"))
        s = s + cur[:, :, -s.shape[-1]:]
    s = F.relu(F.conv1d(s, sd["bottleneck_layer.weight"], b("bottleneck_layer")))
    return F.adaptive_avg_pool1d(s, pool)
