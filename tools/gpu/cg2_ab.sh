#!/bin/bash
# CTA-pair (cta_group::2) N = 256 engine: parity tests, then bench.py with the pairs on and off on the same box
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x -k "gemm or conv or trunk or model or strong or pipeline" 2>&1 | tail -5
for cg in 1 0; do
  AVVAD_CG2=$cg timeout 600 python bench.py > gpurun_out/bench_cg$cg.log 2> gpurun_out/bench_cg$cg.err; echo "bench cg2=$cg exit=$?"
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/bench_cg$cg.log") if x.startswith("{")][-1]
    d = json.loads(l)
    print("CG2=$cg value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"], 4), "clocks", d.get("clocks"))
    print("   breakdown", d.get("breakdown_ms_per_step"))
    print("   train ms", d.get("train", {}).get("ms_per_step"), "parity", d.get("parity", {}).get("ok"))
except Exception as e:
    print("no bench line:", e); import subprocess; print(open("gpurun_out/bench_cg$cg.err").read()[-1500:])
PY
done
