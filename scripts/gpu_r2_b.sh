#!/bin/bash
# round 2: N-GPU visit (N = $1): driver-style bench launch with the sharded leg, NCCL parity test, full compare over N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
grep '^{' gpurun_out/bench_n$N.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e_pageable']['value'])
for s in d.get('sharded') or []: print(json.dumps(s)[:1800])
"
tail -3 gpurun_out/bench_n$N.err
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded_n$N.log 2>&1; echo "pytest sharded rc=$?"; tail -2 gpurun_out/pytest_sharded_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/full_compare.py --log2 26 > gpurun_out/full_compare_2p26_n$N.json 2> gpurun_out/full_compare_n$N.err; echo "full compare rc=$?"
cat gpurun_out/full_compare_2p26_n$N.json; tail -3 gpurun_out/full_compare_n$N.err
