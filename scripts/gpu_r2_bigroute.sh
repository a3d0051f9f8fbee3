#!/bin/bash
# big-ring products through the drop-in symbol: sharded plan on one rank (default) vs layer-per-launch path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "big_ring or odd_sizes or wrapper" > gpurun_out/bigroute_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/bigroute_pytest.log
for v in 1 0; do
MPIRFFT_BIG_PLAN=$v timeout 900 python - > gpurun_out/bigroute_$v.log 2> gpurun_out/bigroute_$v.err <<PY
import json, sys, time, torch
sys.path.insert(0, ".")
import bench, mpir_fft_b200 as M
torch.cuda.set_device(0); M.init(0)
dev = torch.device("cuda", 0)
cases = [(1 << 22, 16, 1), (1 << 24, 16, 1), (1 << 24, 15, 4)] + ([(1 << 25, 16, 2), (1 << 26, 16, 4)] if $v else [])
for n, d, w in cases:
    a = bench.splitmix64_dev(torch, 1, n, dev); b = bench.splitmix64_dev(torch, 2, n, dev)
    r = torch.zeros(2*n, dtype=torch.int64, device=dev)
    pl = M.MulPlan(n, n, d, w)
    ts = []
    for it in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); pl.exec_device(r.data_ptr(), a.data_ptr(), b.data_ptr(), torch.cuda.current_stream().cuda_stream); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    from mpir_fft_b200 import residues as R
    ok = R.product_matches(R.residues(a), R.residues(b), R.residues(r))
    print(json.dumps({"n": n, "depth": d, "w": w, "limbs": pl.params["limbs"] if hasattr(pl, "params") else None, "ms": ts, "exact": bool(ok)}), flush=True)
    del pl
PY
echo "BIG_PLAN=$v rc=$?"; cat gpurun_out/bigroute_$v.log; tail -3 gpurun_out/bigroute_$v.err
done
timeout 300 python - <<PY
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import mpir_fft_b200 as M
M.init(0)
L = M.lib()
for limbs in (4, 64, 1024):
    a = np.random.default_rng(1).integers(0, 1 << 63, limbs + 1, dtype=np.uint64); a[limbs] = 0
    r = np.zeros(limbs + 1, dtype=np.uint64)
    f = L.mpn_mul_2expmod_2expp1
    pa, pr = a.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p)
    for _ in range(200): f(pr, pa, C.c_long(limbs), C.c_ulong(5))
    t = time.perf_counter()
    for _ in range(3000): f(pr, pa, C.c_long(limbs), C.c_ulong(5))
    print("mpn_mul_2expmod_2expp1 limbs=%d: %.1f us per call" % (limbs, (time.perf_counter() - t) / 3000 * 1e6))
PY
echo "percall rc=$?"
