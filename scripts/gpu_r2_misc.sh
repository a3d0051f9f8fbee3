#!/bin/bash
# (1) tile kernel: one warp polls the load mbarrier (default) vs all warps (MPIRFFT_TILE_DEBUG=4); (2) where the
# reference's test_mul_2expmod spends its time against this library
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_misc.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_misc.log
for v in 0 4 0 4; do
for wl in cfg2 cfg1; do
env MPIRFFT_TILE_DEBUG=$v timeout 600 python bench.py --workload $wl --steps 40 --warmup 5 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_${wl}_dbg$v.log 2> gpurun_out/bench_${wl}_dbg$v.err; echo "TILE_DEBUG=$v bench $wl rc=$?"
grep '^{' gpurun_out/bench_${wl}_dbg$v.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print('ms',d['ms_per_step'],'median',d.get('step_ms_median'),'minmax',d.get('step_ms_min_max'),'slowest',d.get('slowest_step_index'),'e2e',d['e2e']['ms_per_step'],'exact',d['bit_exact_vs_gmp'],'frac',r.get('frac'))
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict) and v['launches_per_product']})
"
done; done
timeout 900 python - <<'PY'
import sys, time, subprocess, os
sys.path.insert(0, "tests")
import test_dropin_reference_suite as T
t = time.time()
p = T.start_reference_test(T.OURS, "test_mul_2expmod", 0)
out, err = p.communicate(timeout=800)
print("test_mul_2expmod alone: %.1f s" % (time.time() - t), out.strip()[-200:], err.strip()[-200:])
PY
timeout 300 python - <<'PY'
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import mpir_fft_b200 as M
M.init(0)
L = M.lib()
f = L.mpn_mul_2expmod_2expp1
for limbs in (2, 60, 61, 62, 500, 1953):
    a = np.random.default_rng(1).integers(0, 1 << 63, limbs + 1, dtype=np.uint64); a[limbs] = 0
    r = np.zeros(limbs + 1, dtype=np.uint64)
    pa, pr = a.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p)
    t = time.perf_counter()
    for d in range(64): f(pr, pa, C.c_long(limbs), C.c_ulong(d))
    print("limbs=%d: %.1f us per call (first 64 calls at this size)" % (limbs, (time.perf_counter() - t) / 64 * 1e6))
PY
