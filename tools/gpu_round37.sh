#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_g2.log 2>&1; echo "g2 exit=$?"
grep '^{' gpurun_out/bench_g2.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['variants']['dedup_video']['value'])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_g2.log 2>&1; echo "ref exit=$?"
grep '^{' gpurun_out/bench_ref_g2.log | cut -c1-300
