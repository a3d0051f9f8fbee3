#!/bin/bash
# N-GPU visit (default 8): driver-style bench launch (replicas + the sharded legs: 2^26 with the 1-GPU baseline, and the largest size), NCCL parity test
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
grep '^{' gpurun_out/bench_n$N.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e_pageable']['value'])
for s in d.get('sharded') or []: print(json.dumps(s)[:2200])
"
tail -5 gpurun_out/bench_n$N.err
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded_n$N.log 2>&1; echo "pytest sharded rc=$?"; tail -2 gpurun_out/pytest_sharded_n$N.log
