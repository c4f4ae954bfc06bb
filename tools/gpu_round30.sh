#!/bin/bash
mkdir -p gpurun_out
for cfg in "24576 64"; do
set -- $cfg
AVVAD_CHUNK=$1 AVVAD_PIECE=$2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.log 2>&1; echo "chunk=$1 piece=$2 exit=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_s.log') if x.startswith('{')][-1]
d=json.loads(l)
print(round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['breakdown_ms_per_step']['conv_tc'], d['breakdown_ms_per_step']['stem_tc'], d['roofline']['frac'])
PY
done
