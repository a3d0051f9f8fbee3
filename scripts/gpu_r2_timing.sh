#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
echo "=== MPIRFFT_TILE_DEBUG=$dbg"
MPIRFFT_TILE_DEBUG=$dbg timeout 300 python scripts/tile_timing.py 2>&1 | tee gpurun_out/tile_timing_$dbg.log | grep -v quantiles
done
