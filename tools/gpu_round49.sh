#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "exit=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print(round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['breakdown_ms_per_step'], d['roofline']['frac'], d['variants']['dedup_video']['value'], d['cpu_baseline']['value'])
PY
