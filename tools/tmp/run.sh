#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 200 -x -k "conv or layer1" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -x 2>&1 | tail -2
bash tools/gpu/per_layer.sh 2>&1 | tail -11
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_s.log | head -1
