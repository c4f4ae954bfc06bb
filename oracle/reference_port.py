"""CPU baseline: the reference's AV forward expressed with the same library modules it uses
(torchvision resnet18, nn.LSTM over packed sequences, nn.Linear), loaded from the same state_dict.
Test / benchmark infrastructure only (bench.py's cpu_baseline leg and `--impl reference`).

Follows packages/models/AV_Net.py:12-141; the MCB branch uses torch.fft because the legacy
torch.rfft/irfft calls of packages/models/compact_bilinear_pooling.py:152-171 no longer exist.
The per-utterance front end follows scripts/evaluate_AV_net.py:176-233 (CPU torch.stft)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import frontend as ofe
from . import models as om
from . import video as ov


class RefDeepVADAV(nn.Module):
    def __init__(self, lstm_layers=2, hidden=1024, y_dim=1, use_mcb=True, eps=1e-8):
        super().__init__()
        import torchvision.models as tvm

        self.use_mcb, self.eps = use_mcb, eps
        self.features = nn.Sequential(*list(tvm.resnet18(weights=None).children())[:-1])
        self.bn = nn.BatchNorm1d(512, eps=eps)
        if use_mcb:
            for n, size in (("1", 513), ("2", 512)):
                self.register_buffer(f"mcb_sketch{n}_h", torch.zeros(size, dtype=torch.int64))
                self.register_buffer(f"mcb_sketch{n}_s", torch.ones(size))
            self.mcb_bn = nn.BatchNorm1d(1024, eps=eps)
            in_size = 1024
        else:
            in_size = 1025
        self.lstm_merged = nn.LSTM(in_size, hidden, lstm_layers)
        self.vad_merged = nn.Linear(hidden, y_dim)

    def load_reference_state_dict(self, sd):
        own = {}
        for k, v in sd.items():
            if k.startswith("mcb.sketch"):
                own["mcb_sketch" + k[len("mcb.sketch"):].replace(".", "_")] = v
            else:
                own[k] = v
        self.load_state_dict(own)
        return self

    def forward(self, audio, video, lengths):
        B, T, H, W = video.shape
        v = video.unsqueeze(2).repeat(1, 1, 3, 1, 1).view(B * T, 3, H, W)
        v = self.features(v).squeeze().view(B, T, -1)
        if self.use_mcb:
            sd = {"mcb.sketch1.h": self.mcb_sketch1_h, "mcb.sketch1.s": self.mcb_sketch1_s,
                  "mcb.sketch2.h": self.mcb_sketch2_h, "mcb.sketch2.s": self.mcb_sketch2_s}
            y = om.mcb(audio, v, sd)
            y = torch.sign(y) * torch.sqrt(torch.abs(y) + self.eps)
            y = y / torch.norm(y, p=2).detach()
            y = self.mcb_bn(y.permute(1, 2, 0).contiguous()).permute(2, 0, 1).contiguous()
        else:
            y = torch.cat([audio, v], dim=2)
        y = pack_padded_sequence(y, lengths=lengths, enforce_sorted=False, batch_first=True)
        out, _ = self.lstm_merged(y)
        out, _ = pad_packed_sequence(out, batch_first=True, total_length=T)
        return self.vad_merged(out)


def cpu_av_inputs(waves, videos_u8, audio_mean, audio_std, video_mean, video_std, eps=1e-8):
    """Per-utterance torch.stft front end and frame-rate conversion, collate (zero-pad, THEN standardise:
    packages/utils.py:157-166 -> scripts/train_AV_net.py:287-291).  Returns (audio (B,T,513), video (B,T,67,67), lengths)."""
    feats, vids, lens = [], [], []
    for w, v in zip(waves, videos_u8):
        x = ofe.peak_normalise(w)
        S = ofe.stft_torch_fp32(x)
        lp = np.log(S.real ** 2 + S.imag ** 2 + np.float32(eps)).astype(np.float32)  # (513,T)
        T = min(lp.shape[1], ov.upsampled_length(v.shape[0]))
        feats.append(torch.from_numpy(lp[:, :T].T.copy()))
        vids.append(torch.from_numpy(ov.upsample_gather(v, T)))
        lens.append(T)
    Tm = max(lens)
    B = len(lens)
    a = torch.zeros(B, Tm, 513)
    vv = torch.zeros(B, Tm, 67, 67)
    for i in range(B):
        a[i, : lens[i]] = feats[i]
        vv[i, : lens[i]] = vids[i]
    a = (a - torch.as_tensor(audio_mean).reshape(1, 1, -1)) / (torch.as_tensor(audio_std).reshape(1, 1, -1) + eps)
    vv = (vv - video_mean) / (video_std + eps)
    return a, vv, lens


def cpu_av_step(model: RefDeepVADAV, waves, videos_u8, audio_mean, audio_std, video_mean, video_std, eps=1e-8):
    """One reference-style pass over a list of utterances on the CPU: cpu_av_inputs, batched forward, sigmoid,
    threshold.  Returns (posteriors (B,T), decisions (B,T), lengths)."""
    a, vv, lens = cpu_av_inputs(waves, videos_u8, audio_mean, audio_std, video_mean, video_std, eps)
    with torch.no_grad():
        logits = model(a, vv, lens)[..., 0]
    post = torch.sigmoid(logits)
    return post, (post > 0.5).int(), lens
