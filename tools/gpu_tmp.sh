#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_models.py tests/test_gpu_pipeline.py tests/test_gpu_dataprep.py tests/test_gpu_configs.py -q -m gpu --timeout 200 -x 2>&1 | tail -3
python tools/micro/fft_ab.py
