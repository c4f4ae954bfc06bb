#!/bin/bash
# Round-2 GPU check: parity suite through the C ABI (all failures reported), smoke(), then the benchmark (both arms).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -s 2>&1 | tee gpurun_out/r2_suite.log | grep -v "^$" | tail -${TAIL:-40}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/r2_smoke.log | tail -3
if [ "${BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; echo "bench exit=$?"; tail -5 gpurun_out/r2_bench.err
  python - <<'PY'
import json
try:
    l = [x for x in open("gpurun_out/r2_bench.log") if x.startswith("{")][-1]
    d = json.loads(l)
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["frac"])
    print("breakdown", d.get("breakdown_ms_per_step"))
    print("ragged", d["variants"]["ragged"])
    print("train", d.get("train"))
    print("parity", d.get("parity"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
fi
