#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py -q -m gpu --timeout 200 -x 2>&1 | tail -3
AVVAD_LAYER_DUMP=gpurun_out/stages.json timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print(round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['breakdown_ms_per_step'], d['roofline']['frac'])
s=json.load(open('gpurun_out/stages.json'))
for l in s['layers']: print('  ', l['flops_per_launch'], round(l['ms_total']/s['steps'],3), round(l['tflops'],1))
PY
