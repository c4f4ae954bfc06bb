#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py -q -m gpu --timeout 300 2>&1 | tail -30
