#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_train.py tests/test_gpu_gemm.py -q -m gpu --timeout 200 -x 2>&1 | tail -5
for c in 1 2 4 8; do
AVVAD_LSTM_CLUSTER=$c timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$c.log 2>&1; echo "cluster $c exit=$?"
python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/bench_c$c.log') if x.startswith('{')][-1]
    d=json.loads(l)
    print(round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step']['lstm_step_tc'], d['breakdown_ms_per_step']['gemm_tc'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_c$c.log').read()[-1500:])
PY
done
