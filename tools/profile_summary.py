#!/usr/bin/env python
"""Turns the artefacts of tools/gpu/profile.sh (gpurun_out/launches.csv, gpurun_out/prof_full.ncu-rep, gpurun_out/bench_plain.log)
into the markdown tables committed under profiles/.  Usage: python tools/profile_summary.py > profiles/<name>.md"""
import collections
import csv
import io
import json
import subprocess
import sys

OUT = "gpurun_out"


def launch_table():
    rows = list(csv.reader(open(f"{OUT}/launches.csv")))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    idx = {n: i for i, n in enumerate(hdr)}
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start:]:
        if len(r) < len(hdr):
            continue
        val = float(r[idx["Metric Value"]])
        unit = r[idx["Metric Unit"]]
        val = val / 1000 if unit == "ns" else (val * 1000 if unit == "ms" else val)
        k = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("avvad::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += val
        tot += val
    lines = ["| kernel | launches | ms | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| `{k[:72]}` | {n} | {t / 1000:.3f} | {100 * t / tot:.1f} % |")
    lines.append(f"\ntotal {tot / 1000:.2f} ms in {sum(v[0] for v in agg.values())} launches")
    return "\n".join(lines), agg, tot


def ncu_table(name="prof_full"):
    # tools/gpu/profile.sh exports the raw page on the GPU box (the .ncu-rep itself may exceed gpurun's 64 MiB limit)
    import os
    if os.path.exists(f"{OUT}/{name}_raw.csv") and os.path.getsize(f"{OUT}/{name}_raw.csv") > 0:
        raw = open(f"{OUT}/{name}_raw.csv").read()
    else:
        raw = subprocess.run(["ncu", "-i", f"{OUT}/{name}.ncu-rep", "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h = r[0]
    idx = {n: i for i, n in enumerate(h)}
    want = ["ID", "Kernel Name", "launch__grid_size", "gpu__time_duration.sum",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    units = {w: r[1][idx[w]] for w in want if w in idx}
    lines = ["| id | kernel | grid | us | tensor pipe active % | L2 throughput % | DRAM read MB | DRAM write MB | regs | SM throughput % |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    conv_rd = conv_wr = 0.0
    n_conv = 0

    def mb(v, u):
        v = float(v)
        return v * 1000 if u.lower().startswith("g") else (v / 1000 if u.lower().startswith("k") else v)

    for row in r[2:]:
        g = lambda w: row[idx[w]]
        name = g("Kernel Name").replace("void ", "").replace("avvad::", "").split("(")[0]
        t = float(g("gpu__time_duration.sum"))
        t_us = t * 1000 if units["gpu__time_duration.sum"] == "ms" else (t / 1000 if units["gpu__time_duration.sum"] == "ns" else t)
        rd = mb(g("dram__bytes_read.sum"), units["dram__bytes_read.sum"])
        wr = mb(g("dram__bytes_write.sum"), units["dram__bytes_write.sum"])
        if "tc_slab" in name or "tc_block17" in name or ("tc_tma" in name and rd > 100):  # the xproj GEMMs read only 50 MB
            conv_rd += rd
            conv_wr += wr
            n_conv += 1
        lines.append(f"| {g('ID')} | `{name[:40]}` | {g('launch__grid_size')} | {t_us:.0f} | "
                     f"{float(g('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
                     f"{float(g('lts__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {rd:.0f} | {wr:.0f} | "
                     f"{g('launch__registers_per_thread')} | {float(g('sm__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} |")
    return "\n".join(lines), conv_rd, conv_wr, n_conv


def main():
    lt, agg, tot = launch_table()
    nt, rd, wr, n = ncu_table()
    b = json.loads([x for x in open(f"{OUT}/bench_plain.log") if x.startswith("{")][-1])
    print("## Launch list of the bench command (per-launch times are cold-cache and serialised: compare shares)\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --ncu --warmup 0`, after "
          f"`python bench.py --steps 3 --warmup 3` exited 0 without ncu ({b['ms_per_step']:.1f} ms/step, conv share "
          f"{b['roofline']['share_of_step']:.3f}, {b['roofline']['achieved']:.0f} TFLOP/s on the convolutions).\n")
    print(lt)
    print("\n## `ncu --set full --clock-control none --import-source on`, one pass of 20,288 frames (`--batch 64`)\n")
    print(nt)
    try:
        pt = ncu_table("prof_lstm_pair")[0]
        print("\n## `ncu --set full` of the CTA-pair LSTM recurrence at the benchmark batch (B = 256, T = 317, one launch per layer: "
              "`AVVAD_LSTM_CHUNKS=1 AVVAD_LSTM_COOP=0`)\n")
        print(pt)
    except Exception as e:  # capture absent
        print(f"\n(no LSTM pair capture: {e})")
    # machine-readable copy for bench.py's roofline.traffic (profiles/*traffic*.json, newest file wins)
    if len(sys.argv) > 1:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        frames = 64 * 317
        # algorithmic operand bytes of the same launches: bf16 activations in + out per conv layer (NHWC), weights once
        # layer1: two fused BasicBlocks (x in, z out; conv_a's output never leaves the SM, the residual is x itself)
        layers = [(17, 64, 64, 17)] * 2 + [(17, 64, 128, 9), (9, 128, 128, 9), (9, 128, 128, 9), (9, 128, 128, 9),
                  (9, 128, 256, 5), (5, 256, 256, 5), (5, 256, 256, 5), (5, 256, 256, 5),
                  (5, 256, 512, 3), (3, 512, 512, 3), (3, 512, 512, 3), (3, 512, 512, 3)]
        alg = sum(frames * 2 * (hi * hi * ci + ho * ho * co) for hi, ci, co, ho in layers)
        alg += frames * 2 * (9 * 9 * 128 * 2 + 5 * 5 * 256 * 2 + 3 * 3 * 512 * 2)  # residual reads + ds inputs of layer2-4
        json.dump({"commit": commit, "source": "ncu --set full --clock-control none, bench.py --ncu --warmup 0 --batch 64 "
                                               "(one 20,288-frame trunk pass), dram__bytes_read.sum + dram__bytes_write.sum",
                   "conv_launches": n, "conv_dram_bytes_per_launch": (rd + wr) * 1e6 / max(n, 1),
                   "conv_dram_read_bytes": rd * 1e6, "conv_dram_write_bytes": wr * 1e6,
                   "conv_algorithmic_bytes_per_launch": alg / max(n, 1), "unit": "bytes per launch (average over the "
                   "convolution launches of one 64-utterance pass)"}, open(sys.argv[1], "w"), indent=1)
    print(f"\nDRAM bytes of the {n} convolution launches (slab + TMA launches that read > 100 MB, i.e. without the two xproj GEMMs): {rd / 1000:.2f} GB read + {wr / 1000:.2f} GB written = "
          f"{(rd + wr) / max(n, 1) / 1000:.3f} GB per launch.")


if __name__ == "__main__":
    main()
