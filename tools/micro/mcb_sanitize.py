"""Tiny MCB + front-end call for compute-sanitizer (racecheck / memcheck) of the register-FFT kernels:
   compute-sanitizer --tool racecheck python tools/micro/mcb_sanitize.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "audio-visual-vad_b200"))
import torch  # noqa: E402

from avvad import engine as E  # noqa: E402

torch.manual_seed(0)
rows, T = 48, 12
sd = {"mcb.sketch1.h": torch.randint(0, 1024, (513,)), "mcb.sketch2.h": torch.randint(0, 1024, (512,)),
      "mcb.sketch1.s": torch.randint(0, 2, (513,)).float() * 2 - 1, "mcb.sketch2.s": torch.randint(0, 2, (512,)).float() * 2 - 1,
      "mcb_bn.weight": torch.ones(1024), "mcb_bn.bias": torch.zeros(1024), "mcb_bn.running_mean": torch.zeros(1024),
      "mcb_bn.running_var": torch.ones(1024)}
mcb = E.Mcb()
mcb.load(sd, "cuda")
a = torch.randn(rows, 513, device="cuda")
v = torch.randn(rows, 512, device="cuda").abs()
o = torch.empty(rows, 1024, device="cuda")
mcb.forward(a, v, out_f32=o)
mcb.forward_grouped(a, v, [12, 7, 0, 3], T, out_f32=o)
wave = torch.randn(3, 256 * 11 + 1024, device="cuda") * 0.1
out = torch.empty(3, T, 513, device="cuda")
E.frontend_logpower(wave, [wave.shape[1], 3000, 2000], [12, 9, 5], T, torch.zeros(513, device="cuda"), torch.ones(513, device="cuda"), out=out)
torch.cuda.synchronize()
print("done", float(o.abs().sum()), float(out.abs().sum()))
