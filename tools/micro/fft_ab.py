"""A/B timing of the two FFT-based kernels (audio front end, MCB fusion) at the bench shape: B=256, T=317."""
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "audio-visual-vad_b200"))
from avvad import engine as E  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
B, T = 256, 317
N = (T - 1) * 256 + 1024
wave = torch.randn(B, N, device=dev) * 0.1
ns = [N] * B
nf = [T] * B
mean = torch.zeros(513, device=dev)
std = torch.ones(513, device=dev)
out = torch.empty(B, T, 513, device=dev)
sd = {"mcb.sketch1.h": torch.randint(0, 1024, (513,)), "mcb.sketch2.h": torch.randint(0, 1024, (512,)),
      "mcb.sketch1.s": torch.randint(0, 2, (513,)).float() * 2 - 1, "mcb.sketch2.s": torch.randint(0, 2, (512,)).float() * 2 - 1,
      "mcb_bn.weight": torch.ones(1024), "mcb_bn.bias": torch.zeros(1024), "mcb_bn.running_mean": torch.zeros(1024),
      "mcb_bn.running_var": torch.ones(1024)}
mcb = E.Mcb()
mcb.load(sd, dev)
a = torch.randn(B * T, 513, device=dev)
v = torch.randn(B * T, 512, device=dev)
ob = torch.empty(B * T, 1024, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


print("frontend ms", round(timeit(lambda: E.frontend_logpower(wave, ns, nf, T, mean, std, out=out)), 4))
print("mcb ms", round(timeit(lambda: mcb.forward(a, v, out_bf16=ob)), 4))
