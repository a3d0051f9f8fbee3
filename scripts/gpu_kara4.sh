#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pointwise_modes.py cfg2 order > gpurun_out/pointwise_modes4.log 2>&1; echo "modes rc=$?"
cat gpurun_out/pointwise_modes4.log
timeout 300 python scripts/pointwise_only.py > gpurun_out/mm_route.log 2>&1; echo "mm rc=$?"; cat gpurun_out/mm_route.log
MPIRFFT_MM_INNER=64 timeout 300 python scripts/pointwise_only.py > gpurun_out/mm_route64.log 2>&1; echo "mm64 rc=$?"; cat gpurun_out/mm_route64.log
