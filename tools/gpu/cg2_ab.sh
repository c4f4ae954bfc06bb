#!/bin/bash
# CTA-pair engine A/B: parity tests, then bench.py per configuration on the same box
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x -k "gemm or conv or trunk or model or strong or pipeline" 2>&1 | tail -5
fi
CFGS=${CFGS:-AVVAD_CG2_128=2 AVVAD_CG2_128=1 AVVAD_CG2_128=0}
for cfg in $CFGS; do
  env $cfg timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench_$cfg.log 2> gpurun_out/bench_$cfg.err; echo "bench $cfg exit=$?"
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/bench_$cfg.log") if x.startswith("{")][-1]
    d = json.loads(l)
    print("$cfg value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"], 4), "clocks", d.get("clocks"))
    print("   breakdown", d.get("breakdown_ms_per_step"))
    print("   train ms", d.get("train", {}).get("ms_per_step"), "parity", d.get("parity", {}).get("ok"))
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/bench_$cfg.err").read()[-1500:])
PY
done
