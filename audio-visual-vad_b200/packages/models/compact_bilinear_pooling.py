"""Count-sketch / compact-bilinear-pooling parameter holders (reference:
packages/models/compact_bilinear_pooling.py:59-113,222-263).  The random projections ``h`` (int64 in
[0, output_size)) and ``s`` (+-1) are registered as buffers under the same names, so checkpoints
(`mcb.sketch1.h`, `mcb.sketch1.s`, ...) load unchanged.  The arithmetic (sketch -> FFT circular
convolution -> signed sqrt -> L2 -> BN) is fused inside libavvad (csrc/mcb.cu) and driven from
DeepVAD_AV.forward."""
import torch
import torch.nn as nn


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

    def forward(self, x):
        raise NotImplementedError("CountSketch is evaluated inside the fused MCB kernel (DeepVAD_AV.forward)")


class CompactBilinearPooling(nn.Module):
    def __init__(self, input1_size, input2_size, output_size, h1=None, s1=None, h2=None, s2=None,
                 force_cpu_scatter_add=False):
        super().__init__()
        self.add_module('sketch1', CountSketch(input1_size, output_size, h1, s1))
        self.add_module('sketch2', CountSketch(input2_size, output_size, h2, s2))
        self.output_size = output_size
        self.force_cpu_scatter_add = force_cpu_scatter_add

    def forward(self, x, y=None):
        raise NotImplementedError("CompactBilinearPooling is evaluated inside the fused MCB kernel "
                                  "(DeepVAD_AV.forward)")
