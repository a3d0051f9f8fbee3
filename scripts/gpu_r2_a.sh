#!/bin/bash
# round 2, first GPU visit: parity tests, default bench line (with the sharded leg), full compare at 2^26
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 4000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 900 python scripts/full_compare.py --log2 26 > gpurun_out/full_compare_2p26.json 2> gpurun_out/full_compare.err; echo "full compare rc=$?"
cat gpurun_out/full_compare_2p26.json; tail -3 gpurun_out/full_compare.err
