#!/bin/bash
# big-ring path: NCCL/world-1 parity test, the 2^26 sharded leg with both layer executors, pass structure
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in 0 1; do
MPIRFFT_BIG_INPLACE=$v MPIRFFT_VERBOSE=$v timeout 600 python - > gpurun_out/big_$v.log 2> gpurun_out/big_$v.err <<PY
import json, torch, sys
sys.path.insert(0, ".")
import bench, mpir_fft_b200 as M
torch.cuda.set_device(0); M.init(0)
import torch.distributed as dist
rec = bench.sharded_leg(torch, dist, M, 0, 1, 26, steps=2, peak_gbs=6557.1)
print(json.dumps({k: rec[k] for k in ("ms_per_step", "phases_ms", "bit_exact")}))
rec = bench.sharded_leg(torch, dist, M, 0, 1, 24, steps=2, peak_gbs=6557.1)
print(json.dumps({k: rec[k] for k in ("workload", "ms_per_step", "phases_ms", "bit_exact")}))
PY
echo "BIG_INPLACE=$v rc=$?"; cat gpurun_out/big_$v.log; grep -c "sliced" gpurun_out/big_$v.err; grep "pass \|xform" gpurun_out/big_$v.err | head -60
done
