#!/bin/bash
# What the driver runs at round end, in one go: GPU suite, smoke(), both bench arms.
#   gpurun --timeout 1800 -- 'bash tools/gpu/final_check.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-260
timeout 600 python bench.py 2>/dev/null | tee gpurun_out/bench_final.log | cut -c1-400
