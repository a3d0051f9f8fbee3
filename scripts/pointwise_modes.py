"""A/B of the product kernels inside new_mpn_mul on one B200: schoolbook blocks (mode 3), Karatsuba
blocks (mode 2); MPIRFFT_PW_UNROLL=1|4 selects the step-loop unrolling.  Prints one JSON line per
(workload, mode): pointwise ms per product from the library's per-class CUDA events, the whole
product from CUDA events, and bit-exactness vs GMP.   python scripts/pointwise_modes.py"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpir_fft_b200 as M                                    # noqa: E402
from oracle import loader as oracle                          # noqa: E402  (checker only)
from bench import splitmix64                                 # noqa: E402

M.init(0); L = M.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
CASES = {"l64": (1 << 16, 1 << 16, 12, 1), "l128": (1 << 18, 1 << 18, 13, 1), "l256 (cfg2)": (1 << 20, 1 << 20, 14, 1)}
if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
    CASES = {"l256 (cfg2)": CASES["l256 (cfg2)"]}
if len(sys.argv) > 1 and sys.argv[1] == "others":
    CASES = {"l64": CASES["l64"], "l128": CASES["l128"], "l512 (cfg3)": (3000000, 1700000, 14, 2)}
for name, (n1, n2, depth, w) in CASES.items():
    a_h, b_h = splitmix64(11, n1), splitmix64(12, n2)
    want = oracle.gmp_mul(a_h, b_h)
    a = torch.from_numpy(a_h.view(np.int64)).cuda(); b = torch.from_numpy(b_h.view(np.int64)).cuda()
    r = torch.zeros(n1 + n2, dtype=torch.int64, device="cuda")
    plan = M.MulPlan(n1, n2, depth, w)
    for mode in (3, 2):
        L.mpirfft_set_pointwise_mode(mode)
        for _ in range(3):
            plan.exec_device(r.data_ptr(), a.data_ptr(), b.data_ptr(), None)
        torch.cuda.synchronize()
        ok = bool(np.array_equal(r.cpu().numpy().view(np.uint64), want))
        reps, tot = 20, 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.exec_device(r.data_ptr(), a.data_ptr(), b.data_ptr(), None); e1.record()
            torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        ms = (C.c_double * 6)(); ln = (C.c_uint64 * 6)(); by = (C.c_double * 6)()
        L.mpirfft_profile_enable(1)
        for _ in range(5):
            flush.zero_(); plan.exec_device(r.data_ptr(), a.data_ptr(), b.data_ptr(), None)
        L.mpirfft_profile_read(ms, ln, by, 6); L.mpirfft_profile_enable(0)
        print(json.dumps({"workload": name, "limbs": plan.params["limbs"], "products": plan.params["trunc"],
                          "mode": {3: "schoolbook", 2: "karatsuba"}[mode],
                          "unroll": os.environ.get("MPIRFFT_PW_UNROLL", "4"), "pointwise_ms": ms[2] / 5, "product_ms": tot / reps, "bit_exact_vs_gmp": ok}), flush=True)
    L.mpirfft_set_pointwise_mode(0)
    plan.close()
