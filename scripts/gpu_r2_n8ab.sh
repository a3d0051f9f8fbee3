#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
MPIRFFT_BIG_INPLACE=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$v scripts/leg.py 30 > gpurun_out/leg30_$v.log 2> gpurun_out/leg30_$v.err; echo "BIG_INPLACE=$v rc=$?"
grep '^{' gpurun_out/leg30_$v.log | cut -c1-900; tail -2 gpurun_out/leg30_$v.err | cut -c1-300
done
