#!/bin/bash
# N-GPU visit: sharded 2^26 (and 2^28 if asked) products
N=$1
mkdir -p gpurun_out
run() { timeout $4 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $5 scripts/big_sharded.py --log2 $1 --depth $2 --w $3 --steps 2 --warmup 1 > gpurun_out/big_2p$1_n$N.log 2>&1; echo "2^$1 rc=$?"; grep -h '^{' gpurun_out/big_2p$1_n$N.log | cut -c1-330; tail -2 gpurun_out/big_2p$1_n$N.log | grep -v '^{' | cut -c1-300; }
run 26 17 1 300 29541
if [ "$2" = "28" ]; then run 28 18 1 500 29542; fi
