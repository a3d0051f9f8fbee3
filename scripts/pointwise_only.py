"""Batched mulmod 2^(64 l)+1 on resident operands: the transform route of mm.c (l >= 250, the
reference's FFT_mulmod_2expp1 recursion) against the direct product kernel, per batch.
   python scripts/pointwise_only.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpir_fft_b200 as M                      # noqa: E402

M.init(0); L = M.lib()
for l, cnt in ((256, 16640), (512, 18432)):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2**64, (cnt, l + 1), dtype=np.uint64); a[:, l] = 0
    b = rng.integers(0, 2**64, (cnt, l + 1), dtype=np.uint64); b[:, l] = 0
    da, db = L.mpirfft_malloc_device(a.nbytes), L.mpirfft_malloc_device(b.nbytes)
    L.mpirfft_memcpy_h2d(da, a.ctypes.data, a.nbytes, None); L.mpirfft_memcpy_h2d(db, b.ctypes.data, b.nbytes, None)
    for it in range(4):
        L.mpirfft_stream_sync(None); t = time.perf_counter()
        L.mpirfft_mulmod_batch_device(da, db, cnt, l, l + 1, None)
        L.mpirfft_stream_sync(None)
        print("mulmod batch (mm.c route) %d x l=%d: %.3f ms" % (cnt, l, (time.perf_counter() - t) * 1e3), flush=True)
    L.mpirfft_free_device(da); L.mpirfft_free_device(db)
