#!/bin/bash
# Full GPU check: parity suite through the C ABI, then the benchmark (both arms).  Run with
#   gpurun --timeout 1500 -- 'bash tools/gpu/suite_and_bench.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tee gpurun_out/suite.log | tail -3
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "reference exit=$?"
python - <<'PY'
import json
for f in ("bench", "bench_ref"):
    l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["value"]), d["ms_per_step"], d.get("breakdown_ms_per_step"), (d.get("roofline") or {}).get("frac"))
PY
