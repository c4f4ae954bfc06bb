"""A/B of the FFT-based kernels -- MCB row pass and log-power front end -- on the radix-4 shared-memory FFT vs the
register FFT (AVVAD_MCB_REG / AVVAD_FE_REG = 0/1) at the bench shape B = 256 x T = 317: device time of the whole
avvad_mcb_forward / avvad_frontend_logpower call and agreement of the fp32 outputs.  The kernel choice is read once per
process, so each mode runs in a child process."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "audio-visual-vad_b200"))


def child(mode: str, out_path: str):
    import torch

    from avvad import engine as E

    torch.manual_seed(0)
    dev = "cuda"
    rows = 256 * 317
    sd = {"mcb.sketch1.h": torch.randint(0, 1024, (513,)), "mcb.sketch2.h": torch.randint(0, 1024, (512,)),
          "mcb.sketch1.s": torch.randint(0, 2, (513,)).float() * 2 - 1,
          "mcb.sketch2.s": torch.randint(0, 2, (512,)).float() * 2 - 1,
          "mcb_bn.weight": torch.ones(1024), "mcb_bn.bias": torch.zeros(1024), "mcb_bn.running_mean": torch.zeros(1024),
          "mcb_bn.running_var": torch.full((1024,), 1.0 / (rows * 1024.0))}
    mcb = E.Mcb()
    mcb.load(sd, dev)
    a = torch.randn(rows, 513, device=dev)
    v = torch.randn(rows, 512, device=dev).abs()
    ob = torch.empty(rows, 1024, dtype=torch.bfloat16, device=dev)
    o32 = torch.empty(rows, 1024, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    mcb.forward(a, v, out_bf16=ob, out_f32=o32)
    torch.cuda.synchronize()
    B, T = 256, 317
    N = (T - 1) * 256 + 1024
    wave = torch.randn(B, N, device=dev) * 0.1
    mean = torch.randn(513, device=dev)
    std = torch.rand(513, device=dev) + 1.0
    fe = torch.empty(B, T, 513, device=dev)
    nsamp = [N - 37 * i for i in range(B)]
    nfr = [T - (i % 5) for i in range(B)]
    E.frontend_logpower(wave, nsamp, nfr, T, mean, std, out=fe)
    torch.cuda.synchronize()
    # grouped call (one norm per utterance) on ragged lengths: both row kernels skip the rows behind a length
    gl = [T - 3 * (i % 7) for i in range(B)]
    og = torch.empty(rows, 1024, dtype=torch.float32, device=dev)
    mcb.forward_grouped(a, v, gl, T, out_f32=og)
    torch.cuda.synchronize()
    torch.save((o32[::97].cpu(), fe[::3].cpu(), og[::89].cpu()), out_path)
    ts = []
    for i in range(23):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mcb.forward(a, v, out_bf16=ob)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"AVVAD_MCB_REG={mode}: mcb forward (row + norm + apply) {ts[len(ts) // 2]:.4f} ms median, {ts[0]:.4f} min")
    ts = []
    for i in range(23):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        E.frontend_logpower(wave, nsamp, nfr, T, mean, std, out=fe)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"AVVAD_FE_REG={mode}: front end (peak + log-power frames) {ts[len(ts) // 2]:.4f} ms median, {ts[0]:.4f} min")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1], sys.argv[2])
        sys.exit(0)
    import torch

    outs = []
    for mode in ("0", "1"):
        path = f"/tmp/mcb_ab_{mode}.pt"
        env = dict(os.environ, AVVAD_MCB_REG=mode, AVVAD_FE_REG=mode)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), mode, path], env=env)
        outs.append(torch.load(path))
    for name, a, b in (("mcb", outs[0][0], outs[1][0]), ("front end", outs[0][1], outs[1][1]),
                       ("mcb grouped", outs[0][2], outs[1][2])):
        d = (a - b).abs().max().item()
        rel = ((a - b).norm() / a.norm()).item()
        print(f"{name}: register FFT vs shared-memory FFT: max |diff| {d:.3e}, rel fro {rel:.3e} (output std {a.std().item():.3f})")
        assert rel < 1e-5, (name, rel)
