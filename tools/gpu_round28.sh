#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_train.py tests/test_gpu_pipeline.py -q -s -m gpu --timeout 300 2>&1 | tail -3
for v in 0 8; do
AVVAD_LSTM_VARIANT=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.log 2>&1; echo "variant $v exit=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_v$v.log') if x.startswith('{')][-1]
d=json.loads(l)
print(d['value'], d['ms_per_step'], d['breakdown_ms_per_step']['lstm_step_tc'])
PY
done
