import sys; sys.path.insert(0, '.')
import numpy as np, ctypes as C, time
import mpir_fft_b200 as M
M.init(0); L = M.lib()
l, cnt = 256, 16640
rng = np.random.default_rng(1)
a = rng.integers(0, 2**64, (cnt, l+1), dtype=np.uint64); a[:, l] = 0
b = rng.integers(0, 2**64, (cnt, l+1), dtype=np.uint64); b[:, l] = 0
da, db = L.mpirfft_malloc_device(a.nbytes), L.mpirfft_malloc_device(b.nbytes)
L.mpirfft_memcpy_h2d(da, a.ctypes.data, a.nbytes, None); L.mpirfft_memcpy_h2d(db, b.ctypes.data, b.nbytes, None)
for it in range(3):
    L.mpirfft_stream_sync(None); t = time.perf_counter()
    L.mpirfft_mulmod_batch_device(da, db, cnt, l, l+1, None)
    print("mulmod batch %d x l=%d: %.3f ms" % (cnt, l, (time.perf_counter()-t)*1e3))
