#!/bin/bash
# A/B of the pointwise kernel's occupancy: 8 warps per SM (240/252 registers) vs 9 (224, MPIRFFT_PW_WARPS=9)
# (record of a finished experiment: the MPIRFFT_PW_WARPS switch it drives was removed after this run -- nine warps per SM were 36 % slower)
mkdir -p gpurun_out
MPIRFFT_PW_WARPS=9 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mulmod or cfg or odd_sizes or mul6" > gpurun_out/pytest_pw9.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_pw9.log
for v in 8 9; do
for wl in cfg2 cfg3 cfg1 big; do
env MPIRFFT_PW_WARPS=$v timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_${wl}_pw$v.log 2> gpurun_out/bench_${wl}_pw$v.err; echo "PW_WARPS=$v bench $wl rc=$?"
grep '^{' gpurun_out/bench_${wl}_pw$v.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'exact',d['bit_exact_vs_gmp'], 'lib', (d.get('library_parameter_choice') or {}).get('ms_per_step'))
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict) and v['launches_per_product']})
"
tail -2 gpurun_out/bench_${wl}_pw$v.err
done; done
for v in 8 9; do
MPIRFFT_PW_WARPS=$v timeout 300 python - <<PY
import json, torch, sys
sys.path.insert(0, ".")
import bench, mpir_fft_b200 as M
torch.cuda.set_device(0); M.init(0)
import torch.distributed as dist
rec = bench.sharded_leg(torch, dist, M, 0, 1, 26, steps=2, peak_gbs=6557.1)
print("PW_WARPS=$v", json.dumps({k: rec[k] for k in ("workload", "ms_per_step", "phases_ms", "bit_exact")}))
PY
done
