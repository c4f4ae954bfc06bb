#!/bin/bash
mkdir -p gpurun_out
AVVAD_LSTM_TRACE=gpurun_out/lstm_trace.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tr.log 2>&1; echo "exit=$?"
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print(round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step']['lstm_step_tc'])
PY
