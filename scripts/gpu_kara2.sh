#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/pointwise_modes2.log
for u in 1 2 4 8; do
MPIRFFT_PW_UNROLL=$u timeout 200 python scripts/pointwise_modes.py cfg2 >> gpurun_out/pointwise_modes2.log 2>&1; echo "modes unroll $u rc=$?"
done
cat gpurun_out/pointwise_modes2.log
