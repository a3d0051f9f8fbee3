#!/bin/bash
# GPU visit: parity tests, default bench, cfg4 bench, cfg1/cfg3 bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.log 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"
tail -c 1500 gpurun_out/bench_cfg4.err
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/bench_cfg1.log 2>&1; echo "cfg1 rc=$?"
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 > gpurun_out/bench_cfg3.log 2>&1; echo "cfg3 rc=$?"
timeout 300 python bench.py --workload big --sharded --steps 5 --warmup 3 > gpurun_out/bench_big1.log 2>&1; echo "big rc=$?"
