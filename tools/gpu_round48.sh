#!/bin/bash
mkdir -p gpurun_out
for v in 0 16 0 16; do
AVVAD_LSTM_VARIANT=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print($v, round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step']['lstm_step_tc'], d['breakdown_ms_per_step']['conv_tc'])
PY
done
