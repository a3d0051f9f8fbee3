#!/bin/bash
# multi-GPU visit: sharded NCCL parity test, then the sharded `big` product and the replica bench at N ranks
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded_n$N.log 2>&1; tail -2 gpurun_out/pytest_sharded_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload big --sharded --steps 10 --warmup 3 > gpurun_out/bench_big_sharded_n$N.log 2>&1; echo "sharded rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_n$N.log 2>&1; echo "replicas rc=$?"
grep -h '^{' gpurun_out/bench_big_sharded_n$N.log gpurun_out/bench_cfg2_n$N.log | cut -c1-260
