#!/bin/bash
# Per-stage and per-layer convolution throughput (CUDA events inside bench.py).
#   gpurun --timeout 900 -- 'bash tools/gpu/per_layer.sh'
mkdir -p gpurun_out
AVVAD_LAYER_DUMP=gpurun_out/stages.json timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.log 2>&1
AVVAD_PROFILE_PER_LAUNCH=1 AVVAD_LAYER_DUMP=gpurun_out/layers.json timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l.log 2>&1
python - <<'PY'
import json
for f in ("stages", "layers"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f)
    for l in d["layers"]:
        print("  ", l["flops_per_launch"], l["launches"], round(l["ms_total"] / d["steps"], 3), "ms/step", round(l["tflops"], 1), "TFLOP/s")
PY
