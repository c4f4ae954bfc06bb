"""Count sketch / compact bilinear pooling (reference: packages/models/compact_bilinear_pooling.py:59-113,222-263).

The random projections ``h`` (int64 in [0, output_size)) and ``s`` (+-1) are registered as buffers under the same names,
so checkpoints (`mcb.sketch1.h`, `mcb.sketch1.s`, ...) load unchanged.  Inside DeepVAD_AV the whole fusion (sketch -> FFT
circular convolution -> signed sqrt -> L2 -> BN) is one fused libavvad call (csrc/mcb.cu); used stand-alone, as the
reference allows, both modules run their own libavvad kernels and are differentiable like the reference's
CountSketchFn / CompactBilinearPoolingFn (forward :7-27,140-173, hand-written backward :29-41,175-220)."""
import os
import sys

import torch
import torch.nn as nn

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)


def _tables(h: torch.Tensor, s: torch.Tensor, out_size: int, device):
    """CSR of the inverse map j -> {i : h_i = j} with ascending i inside a bucket (the order the reference's CPU
    scatter_add_ visits colliding inputs), plus int32 h and fp32 s on `device`."""
    hc = h.detach().to("cpu", torch.int64)
    if hc.numel() and (int(hc.min()) < 0 or int(hc.max()) >= out_size):
        raise ValueError("count sketch index out of range [0, output_size)")
    order = torch.sort(hc, stable=True).indices
    counts = torch.bincount(hc, minlength=out_size)
    off = torch.zeros(out_size + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(counts, 0)
    return (off.to(device, torch.int32), order.to(device, torch.int32), hc.to(device, torch.int32),
            s.detach().to(device, torch.float32).contiguous())


class _SketchTables:
    """Per-device cache of the tables of one CountSketch module, refreshed when the buffers change."""

    def __init__(self):
        self.cache = {}

    def get(self, mod: "CountSketch", device):
        key = (device.type, device.index)
        sig = (mod.h.data_ptr(), mod.h._version, mod.s.data_ptr(), mod.s._version)
        hit = self.cache.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, _tables(mod.h, mod.s, mod.output_size, device))
            self.cache[key] = hit
        return hit[1]


class _CountSketchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tables, in_size, out_size):
        from avvad import lib as L
        L.require_cuda(x)
        off, idx, h32, s = tables
        x2 = x.detach().to(torch.float32).reshape(-1, in_size).contiguous()
        out = torch.empty(x2.shape[0], out_size, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            if x2.shape[0]:
                L.check(L.lib().avvad_count_sketch_forward(L.ptr(x2), x2.shape[0], in_size, out_size, L.ptr(off),
                                                           L.ptr(idx), L.ptr(s), L.ptr(out), L.stream_ptr()))
        ctx.tables, ctx.sizes, ctx.shape = tables, (in_size, out_size), x.shape
        return out.view(x.shape[:-1] + (out_size,))

    @staticmethod
    def backward(ctx, go):
        from avvad import lib as L
        off, idx, h32, s = ctx.tables
        in_size, out_size = ctx.sizes
        g2 = go.detach().to(torch.float32).reshape(-1, out_size).contiguous()
        gx = torch.empty(g2.shape[0], in_size, dtype=torch.float32, device=go.device)
        with torch.cuda.device(go.device):
            if g2.shape[0]:
                L.check(L.lib().avvad_count_sketch_backward(L.ptr(g2), g2.shape[0], in_size, out_size, L.ptr(h32),
                                                            L.ptr(s), L.ptr(gx), L.stream_ptr()))
        return gx.view(ctx.shape), None, None, None


class _McbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, t1, t2):
        from avvad import lib as L
        L.require_cuda(x, y)
        x2 = x.detach().to(torch.float32).reshape(-1, 513).contiguous()
        y2 = y.detach().to(torch.float32).reshape(-1, 512).contiguous()
        rows = x2.shape[0]
        out = torch.empty(rows, 1024, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            if rows:
                L.check(L.lib().avvad_mcb_raw_forward(L.ptr(t1[0]), L.ptr(t1[1]), L.ptr(t1[3]), L.ptr(t2[0]), L.ptr(t2[1]),
                                                      L.ptr(t2[3]), L.ptr(x2), L.ptr(y2), rows, L.ptr(out),
                                                      L.stream_ptr()))
        ctx.save_for_backward(x2, y2)
        ctx.t1, ctx.t2, ctx.xs, ctx.ys = t1, t2, x.shape, y.shape
        return out.view(x.shape[:-1] + (1024,))

    @staticmethod
    def backward(ctx, go):
        from avvad import lib as L
        x2, y2 = ctx.saved_tensors
        t1, t2 = ctx.t1, ctx.t2
        rows = x2.shape[0]
        g2 = go.detach().to(torch.float32).reshape(rows, 1024).contiguous()
        gx = torch.empty(rows, 513, dtype=torch.float32, device=go.device) if ctx.needs_input_grad[0] else None
        gy = torch.empty(rows, 512, dtype=torch.float32, device=go.device) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(go.device):
            if rows and (gx is not None or gy is not None):
                L.check(L.lib().avvad_mcb_raw_backward(L.ptr(t1[0]), L.ptr(t1[1]), L.ptr(t1[3]), L.ptr(t1[2]),
                                                       L.ptr(t2[0]), L.ptr(t2[1]), L.ptr(t2[3]), L.ptr(t2[2]), L.ptr(x2),
                                                       L.ptr(y2), L.ptr(g2), rows, L.ptr(gx), L.ptr(gy), L.stream_ptr()))
        return (gx.view(ctx.xs) if gx is not None else None, gy.view(ctx.ys) if gy is not None else None, None, None)


class CountSketch(nn.Module):
    def __init__(self, input_size, output_size, h=None, s=None):
        super().__init__()
        self.input_size = input_size
        self.output_size = output_size
        if h is None:
            h = torch.LongTensor(input_size).random_(0, output_size)
        if s is None:
            s = 2 * torch.Tensor(input_size).random_(0, 2) - 1
        # nn.Module.float()/.double() only cast floating-point tensors, so the int64 `h` stays integral
        # (the reference monkey-patches h.float/h.double for the same purpose, which breaks pickling)
        self.register_buffer('h', h)
        self.register_buffer('s', s)
        object.__setattr__(self, "_tables", _SketchTables())

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_tables", None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        object.__setattr__(self, "_tables", _SketchTables())

    def tables(self, device):
        return self._tables.get(self, torch.device(device))

    def forward(self, x):
        """x (..., input_size) -> (..., output_size) on libavvad (avvad_count_sketch_forward / _backward)."""
        assert x.size(-1) == self.input_size
        return _CountSketchFn.apply(x, self.tables(x.device), self.input_size, self.output_size)


class CompactBilinearPooling(nn.Module):
    def __init__(self, input1_size, input2_size, output_size, h1=None, s1=None, h2=None, s2=None,
                 force_cpu_scatter_add=False):
        super().__init__()
        self.add_module('sketch1', CountSketch(input1_size, output_size, h1, s1))
        self.add_module('sketch2', CountSketch(input2_size, output_size, h2, s2))
        self.output_size = output_size
        self.force_cpu_scatter_add = force_cpu_scatter_add

    def forward(self, x, y=None):
        """out = sketch1(x) (*) sketch2(y), circular convolution over the last axis (reference :140-173)."""
        if y is None:
            y = x
        sizes = (self.sketch1.input_size, self.sketch2.input_size, self.output_size)
        if sizes != (513, 512, 1024):
            raise NotImplementedError(f"libavvad's MCB kernel is built for the model's sizes (513, 512 -> 1024), got "
                                      f"{sizes}; there is no CPU / library fallback")
        assert x.size(-1) == 513 and y.size(-1) == 512
        return _McbFn.apply(x, y, self.sketch1.tables(x.device), self.sketch2.tables(x.device))
