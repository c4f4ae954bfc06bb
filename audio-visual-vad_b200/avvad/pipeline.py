"""End-to-end AV-VAD inference pipeline: raw 16 kHz waveforms + 30 fps mouth-ROI frames in, frame
posteriors / decisions out, everything between on the GPU (SURVEY §8a rows A1-A4, U, V1-V2, F1-F3,
R1, H1-H2; §8f row 2 "on-device batching front end").

This is the public call a user makes for batched inference (the reference has only a B=1,
CPU-STFT evaluation loop: scripts/evaluate_AV_net.py:141-250); the reference-compatible modules in
``packages.models`` share the same engines.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import engine as E
from . import lib as L


class AVVADPipeline:
    def __init__(self, state_dict: Dict[str, torch.Tensor], audio_mean, audio_std, video_mean: float,
                 video_std: float, use_mcb=True, y_dim=1, hidden=1024, layers=2, eps=1e-8, device="cuda"):
        L.require_cuda()
        self.device = torch.device(device)
        self.use_mcb, self.eps, self.y_dim = use_mcb, float(eps), y_dim
        self.audio_mean = torch.as_tensor(audio_mean, dtype=torch.float32).reshape(-1).to(self.device)
        self.audio_std = torch.as_tensor(audio_std, dtype=torch.float32).reshape(-1).to(self.device)
        self.video_mean, self.video_std = float(video_mean), float(video_std)
        with torch.cuda.device(self.device):
            self.trunk = E.ResNet18Trunk()
            self.trunk.load(state_dict, self.device)
            in_size = 1024 if use_mcb else 1025
            self.lstm = E.Lstm(layers, in_size, hidden, y_dim)
            self.lstm.load(state_dict, self.device, "lstm_merged", "vad_merged")
            self.mcb = None
            if use_mcb:
                self.mcb = E.Mcb()
                self.mcb.load(state_dict, self.device, eps)
        self._bufs = {}

    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    @staticmethod
    def frame_counts(n_samples: Sequence[int], n_src: Sequence[int]):
        """T_b = min(STFT frames, upsampled video frames) (packages/data_handling.py:483-486)."""
        return [min(E.stft_num_frames(n), E.upsampled_length(f)) for n, f in zip(n_samples, n_src)]

    def infer_device(self, wave: torch.Tensor, n_samples, video_u8: torch.Tensor, n_src, lengths=None,
                     t_max: Optional[int] = None):
        """wave (B,N) f32 and video (B,F,67,67) u8/f32 already on the device.
        Returns (logits, posteriors, decisions), each (B,t_max,y_dim) on the device."""
        B = wave.shape[0]
        if lengths is None:
            lengths = self.frame_counts(n_samples, n_src)
        if t_max is None:
            t_max = max(lengths)
        dev = self.device
        lens = E._i32(lengths, dev)
        audio = self._buf("audio", (B, t_max, 513), torch.float32)
        E.frontend_logpower(wave, n_samples, lens, t_max, self.audio_mean, self.audio_std, self.eps, True, out=audio)
        frames = self._buf("frames", (B, t_max, 67, 67), torch.float32)
        E.upsample_gather(video_u8, n_src, lens, t_max, self.video_mean, self.video_std, self.eps, True, out=frames)
        M = B * t_max
        x = self._buf("x", (B, t_max, self.lstm.ld), torch.bfloat16)
        xv = x.view(M, self.lstm.ld)
        if self.use_mcb:
            feat = self._buf("feat", (M, 512), torch.float32)
            self.trunk.forward(frames.view(M, 67, 67), feat=feat)
            self.mcb.forward(audio.view(M, 513), feat, out_bf16=xv)
        else:
            x.zero_()
            E.pack_rows_bf16(audio.view(M, 513), xv, 0, False)
            self.trunk.forward(frames.view(M, 67, 67), feat_bf16=xv, col_off=513, want_f32=False)
        logits, post, dec, _ = self.lstm.forward(x, lens, want_post=True, want_dec=True)
        return logits, post, dec

    def infer_host(self, wave_pinned: torch.Tensor, n_samples, video_pinned: torch.Tensor, n_src):
        """Host (pinned) buffers in, host posteriors/decisions out: H2D copies, the whole device path and
        the D2H read-back are enqueued on the current stream (this is what bench.py's `e2e` times)."""
        w = self._buf("wave_dev", tuple(wave_pinned.shape), wave_pinned.dtype)
        v = self._buf("video_dev", tuple(video_pinned.shape), video_pinned.dtype)
        w.copy_(wave_pinned, non_blocking=True)
        v.copy_(video_pinned, non_blocking=True)
        _, post, dec = self.infer_device(w, n_samples, v, n_src)
        hp = self._bufs.get("post_host")
        if hp is None or hp.shape != post.shape:
            hp = torch.empty(post.shape, dtype=post.dtype, pin_memory=True)
            hd = torch.empty(dec.shape, dtype=dec.dtype, pin_memory=True)
            self._bufs["post_host"], self._bufs["dec_host"] = hp, hd
        hd = self._bufs["dec_host"]
        hp.copy_(post, non_blocking=True)
        hd.copy_(dec, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return hp, hd
