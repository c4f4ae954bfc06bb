#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1; do
AVVAD_EPI_DEBUG=$dbg AVVAD_LAYER_DUMP=gpurun_out/stages_d$dbg.json timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d$dbg.log 2>&1; echo "debug $dbg exit=$?"
python - <<PY
import json
s=json.load(open('gpurun_out/stages_d$dbg.json'))
for l in s['layers']: print('  ', l['flops_per_launch'], round(l['ms_total']/s['steps'],3), round(l['tflops'],1))
PY
done
