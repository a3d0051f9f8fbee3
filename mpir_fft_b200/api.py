"""numpy-level mirror of the reference interface for the multiplication path.

Function names and argument meaning follow /root/reference/mul_fft.c; arrays are numpy uint64
(little-endian limbs, mp_limb_t).  Everything computes on the GPU through the C ABI; there is
no CPU path here -- without a CUDA device these functions raise / the library aborts.
"""
import ctypes as C
import numpy as np
from ._lib import lib, MulParams


def _ptr(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def have_gpu():
    return lib().mpirfft_device_count() > 0


def init(device=0):
    rc = lib().mpirfft_init(int(device))
    if rc != 0:
        raise RuntimeError("mpirfft_init(%d) failed: %s" % (device, lib().mpirfft_last_error().decode()))


def mul_params(n1, n2, depth, w):
    """Derived sizes of new_mpn_mul (mul_fft.c:3193-3203); raises ValueError if illegal."""
    p = MulParams()
    if lib().mpirfft_mul_params_get(C.byref(p), n1, n2, depth, w) != 0:
        raise ValueError("illegal new_mpn_mul parameters n1=%d n2=%d depth=%d w=%d" % (n1, n2, depth, w))
    return {k: int(getattr(p, k)) for k, _ in MulParams._fields_}


def choose_params6(n1, n2):
    """(depth, w, sqrt2) of mpirfft_choose_params6: what mpirfft_mpn_mul uses"""
    d, w, sq = C.c_uint64(), C.c_uint64(), C.c_int()
    if lib().mpirfft_choose_params6(n1, n2, C.byref(d), C.byref(w), C.byref(sq)) != 0:
        raise ValueError("no legal (depth, w) for %d x %d limbs" % (n1, n2))
    return int(d.value), int(w.value), bool(sq.value)


def choose_params(n1, n2):
    d, w = C.c_uint64(), C.c_uint64()
    if lib().mpirfft_choose_params(n1, n2, C.byref(d), C.byref(w)) != 0:
        raise ValueError("no legal (depth, w) for %d x %d limbs" % (n1, n2))
    return int(d.value), int(w.value)


def new_mpn_mul(i1, i2, depth, w):
    """r = i1 * i2 (len(i1)+len(i2) limbs) -- the drop-in symbol new_mpn_mul (mul_fft.c:3190)."""
    i1 = np.ascontiguousarray(i1, dtype=np.uint64)
    i2 = np.ascontiguousarray(i2, dtype=np.uint64)
    mul_params(len(i1), len(i2), depth, w)          # raise instead of letting the C symbol abort
    if not have_gpu():
        raise RuntimeError("no CUDA device: mpir_fft_b200 has no CPU path")
    r = np.zeros(len(i1) + len(i2), dtype=np.uint64)
    lib().new_mpn_mul(_ptr(r), _ptr(i1), len(i1), _ptr(i2), len(i2), depth, w)
    return r


def new_mpn_mul6(i1, i2, depth, w):
    """r = i1 * i2 through the sqrt2 transforms -- the drop-in symbol new_mpn_mul6 (mul_fft.c:3573)."""
    i1 = np.ascontiguousarray(i1, dtype=np.uint64)
    i2 = np.ascontiguousarray(i2, dtype=np.uint64)
    p = MulParams()
    if lib().mpirfft_mul6_params_get(C.byref(p), len(i1), len(i2), depth, w) != 0:
        raise ValueError("illegal new_mpn_mul6 parameters n1=%d n2=%d depth=%d w=%d" % (len(i1), len(i2), depth, w))
    if not have_gpu():
        raise RuntimeError("no CUDA device: mpir_fft_b200 has no CPU path")
    r = np.zeros(len(i1) + len(i2), dtype=np.uint64)
    lib().new_mpn_mul6(_ptr(r), _ptr(i1), len(i1), _ptr(i2), len(i2), depth, w)
    return r


def mpn_mul(i1, i2):
    """r = i1 * i2 with automatically chosen (depth, w) -- mpirfft_mpn_mul"""
    i1 = np.ascontiguousarray(i1, dtype=np.uint64)
    i2 = np.ascontiguousarray(i2, dtype=np.uint64)
    choose_params(len(i1), len(i2))                  # raise instead of letting the C symbol abort
    if not have_gpu():
        raise RuntimeError("no CUDA device: mpir_fft_b200 has no CPU path")
    r = np.zeros(len(i1) + len(i2), dtype=np.uint64)
    lib().mpirfft_mpn_mul(_ptr(r), _ptr(i1), len(i1), _ptr(i2), len(i2))
    return r


class MulPlan:
    """Device-resident multiplication plan (mpirfft_mul_plan_*)."""

    def __init__(self, n1, n2, depth, w, sqrt2=False):
        """sqrt2: the plan of new_mpn_mul6 (transform length 4n, mul_fft.c:3573)"""
        self.n1, self.n2, self.depth, self.w, self.sqrt2 = n1, n2, depth, w, bool(sqrt2)
        if sqrt2:
            p = MulParams()
            if lib().mpirfft_mul6_params_get(C.byref(p), n1, n2, depth, w) != 0:
                raise ValueError("illegal new_mpn_mul6 parameters n1=%d n2=%d depth=%d w=%d" % (n1, n2, depth, w))
            self.params = {k: int(getattr(p, k)) for k, _ in MulParams._fields_}
        else:
            self.params = mul_params(n1, n2, depth, w)
        h = C.c_void_p()
        create = lib().mpirfft_mul6_plan_create if sqrt2 else lib().mpirfft_mul_plan_create
        rc = create(C.byref(h), n1, n2, depth, w)
        if rc != 0:
            raise RuntimeError("mpirfft_mul%s_plan_create failed (%d): %s" % ("6" if sqrt2 else "", rc, lib().mpirfft_last_error().decode()))
        self.h = h

    def close(self):
        if self.h:
            lib().mpirfft_mul_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exec_device(self, d_r, d_i1, d_i2, stream=None):
        rc = lib().mpirfft_mul_exec_device(self.h, d_r, d_i1, d_i2, stream)
        if rc != 0:
            raise RuntimeError("mpirfft_mul_exec_device failed (%d): %s" % (rc, lib().mpirfft_last_error().decode()))

    def exec_phase(self, phase, d_r, d_i1, d_i2, stream=None):
        rc = lib().mpirfft_mul_exec_phase(self.h, phase, d_r, d_i1, d_i2, stream)
        if rc != 0:
            raise RuntimeError("mpirfft_mul_exec_phase failed (%d): %s" % (rc, lib().mpirfft_last_error().decode()))

    def exec_host(self, r, i1, i2):
        rc = lib().mpirfft_mul_exec_host(self.h, _ptr(r), _ptr(i1), _ptr(i2))
        if rc != 0:
            raise RuntimeError("mpirfft_mul_exec_host failed (%d): %s" % (rc, lib().mpirfft_last_error().decode()))

    @property
    def launches(self):
        return int(lib().mpirfft_mul_plan_launches(self.h))

    @property
    def device_bytes(self):
        return int(lib().mpirfft_mul_plan_device_bytes(self.h))


def mulmod_batch(a, b):
    """a[k] * b[k] mod 2^(64 l)+1 for blocks a[k], b[k] of l+1 limbs (canonical). Returns blocks."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    assert a.shape == b.shape and a.ndim == 2
    count, size = a.shape
    L = lib()
    if not have_gpu():
        raise RuntimeError("no CUDA device: mpir_fft_b200 has no CPU path")
    da, db = L.mpirfft_malloc_device(a.nbytes), L.mpirfft_malloc_device(b.nbytes)
    if not da or not db:
        raise MemoryError("device allocation failed")
    try:
        L.mpirfft_memcpy_h2d(da, _ptr(a), a.nbytes, None)
        L.mpirfft_memcpy_h2d(db, _ptr(b), b.nbytes, None)
        rc = L.mpirfft_mulmod_batch_device(da, db, count, size - 1, size, None)
        if rc != 0:
            raise RuntimeError("mpirfft_mulmod_batch_device failed (%d): %s" % (rc, L.mpirfft_last_error().decode()))
        out = np.empty_like(a)
        L.mpirfft_memcpy_d2h(_ptr(out), da, a.nbytes, None)
        L.mpirfft_stream_sync(None)
    finally:
        L.mpirfft_free_device(da)
        L.mpirfft_free_device(db)
    return out
