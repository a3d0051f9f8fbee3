#!/bin/bash
# the clocks key of the bench line: NVML polling, and the nvidia-smi fallbacks when NVML is unavailable
mkdir -p gpurun_out
for v in "" 1; do
MPIRFFT_BENCH_NO_NVML=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_clk$v.log 2> gpurun_out/bench_clk$v.err; echo "NO_NVML='$v' rc=$?"
grep '^{' gpurun_out/bench_clk$v.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'])"
tail -2 gpurun_out/bench_clk$v.err
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/bench.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_ms_median'], d['e2e']['ms_per_step'], d['e2e_pageable']['ms_per_step'], d['clocks'], d['roofline']['frac'])"
