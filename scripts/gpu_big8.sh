#!/bin/bash
# 8-GPU visit: one product of 2^26, 2^28 and 2^30 limbs per operand, sharded over the box
mkdir -p gpurun_out
run() { timeout $4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $5 scripts/big_sharded.py --log2 $1 --depth $2 --w $3 --steps 2 --warmup 1 > gpurun_out/big_2p$1_n8.log 2>&1; echo "2^$1 rc=$?"; grep -h '^{' gpurun_out/big_2p$1_n8.log | cut -c1-420; tail -3 gpurun_out/big_2p$1_n8.log | grep -v '^{' | cut -c1-300; }
run 26 17 1 300 29531
run 28 18 1 400 29532
run 30 19 1 600 29533
