#!/bin/bash
# One GPU-box session: runs the parity suites with per-suite timeouts and keeps every log under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 15 gpurun_out/$name.log
}
run gemm 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --timeout 120
run gemm_bn64 300 env AVVAD_BN=64 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120
run gemm_bn256 300 env AVVAD_BN=256 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120
run frontend 300 python -m pytest tests/test_gpu_frontend.py -q -m gpu --timeout 120
run models 600 python -m pytest tests/test_gpu_models.py -q -m gpu -s --timeout 300
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
