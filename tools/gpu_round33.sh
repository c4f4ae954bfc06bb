#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -2
bash tools/gpu_round31.sh
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_s.log') if x.startswith('{')][-1]
d=json.loads(l)
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step']); print(d['breakdown_ms_per_step']); print(d['roofline']['frac'])
PY
