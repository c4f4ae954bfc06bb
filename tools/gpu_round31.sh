#!/bin/bash
mkdir -p gpurun_out
AVVAD_LAYER_DUMP=gpurun_out/stages.json timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.log 2>&1; echo "exit=$?"
cat gpurun_out/stages.json | python -c "
import json,sys
d=json.load(sys.stdin)
for l in d['layers']: print(l['flops_per_launch'], l['launches'], round(l['ms_total']/d['steps'],3), round(l['tflops'],1))
"
AVVAD_PROFILE_PER_LAUNCH=1 AVVAD_LAYER_DUMP=gpurun_out/layers.json timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l.log 2>&1; echo "exit=$?"
cat gpurun_out/layers.json | python -c "
import json,sys
d=json.load(sys.stdin)
for l in d['layers']: print(l['flops_per_launch'], l['launches'], round(l['ms_total']/d['steps'],3), round(l['tflops'],1))
"
