#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_wavenet.py -q -m gpu --timeout 200 -x 2>&1 | tail -5
for mb in 2 1; do
AVVAD_MB=$mb AVVAD_LAYER_DUMP=gpurun_out/stages_mb$mb.json timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb$mb.log 2>&1; echo "mb $mb exit=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_mb$mb.log') if x.startswith('{')][-1]
d=json.loads(l)
print(round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step']['conv_tc'], d['roofline']['frac'])
s=json.load(open('gpurun_out/stages_mb$mb.json'))
for l in s['layers']: print('  ', l['flops_per_launch'], round(l['ms_total']/s['steps'],3), round(l['tflops'],1))
PY
done
