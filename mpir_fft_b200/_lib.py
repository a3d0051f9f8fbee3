"""ctypes loader for libmpirfft_b200.so (the C-ABI shared library built in-tree).

There is deliberately no fallback: if the library is missing it is built (nvcc, sm_100a); if
that fails, or if a compute entry point is called without a CUDA device, an error is raised.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpirfft_b200.so")
CSRC = os.path.join(_HERE, "csrc")

u64p = C.POINTER(C.c_uint64)
u64pp = C.POINTER(u64p)


class MulParams(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in
                ("n", "bits1", "sqrt", "j1", "j2", "trunc", "limbs", "n2", "trunc_rows")]


def build(force=False):
    """Compile the library in-tree (plain C host + sm_100a kernels)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", CSRC, "clean"])
    subprocess.check_call(["make", "-s", "-C", CSRC])
    return LIB_PATH


_lib = None


def lib():
    """The product library (always mpir_fft_b200/libmpirfft_b200.so, never anything else)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = bind(C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL))
    return _lib


def bind(L):
    """Attach the C-ABI signatures of include/mpirfft_b200.h to a loaded library handle."""
    vp, sz, u64, i64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_long, C.c_int
    sig = {
        "mpirfft_init": (i32, [i32]),
        "mpirfft_device_count": (i32, []),
        "mpirfft_last_error": (C.c_char_p, []),
        "mpirfft_version": (C.c_char_p, []),
        "mpirfft_mul_params_get": (i32, [C.POINTER(MulParams), i64, i64, u64, u64]),
        "mpirfft_mul6_params_get": (i32, [C.POINTER(MulParams), i64, i64, u64, u64]),
        "mpirfft_mul6_plan_create": (i32, [C.POINTER(vp), i64, i64, u64, u64]),
        "new_mpn_mul6": (None, [vp, vp, i64, vp, i64, u64, u64]),
        "mpirfft_choose_params": (i32, [i64, i64, C.POINTER(u64), C.POINTER(u64)]),
        "mpirfft_choose_params6": (i32, [i64, i64, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]),
        "mpirfft_mul_plan_create": (i32, [C.POINTER(vp), i64, i64, u64, u64]),
        "mpirfft_mul_plan_destroy": (None, [vp]),
        "mpirfft_mul_exec_device": (i32, [vp, vp, vp, vp, vp]),
        "mpirfft_mul_exec_host": (i32, [vp, vp, vp, vp]),
        "mpirfft_mul_exec_phase": (i32, [vp, i32, vp, vp, vp, vp]),
        "mpirfft_mul_plan_device_bytes": (sz, [vp]),
        "mpirfft_mul_plan_launches": (u64, [vp]),
        "mpirfft_mulmod_batch_device": (i32, [vp, vp, sz, sz, sz, vp]),
        "mpirfft_malloc_device": (vp, [sz]),
        "mpirfft_free_device": (None, [vp]),
        "mpirfft_malloc_pinned": (vp, [sz]),
        "mpirfft_free_pinned": (None, [vp]),
        "mpirfft_memcpy_h2d": (i32, [vp, vp, sz, vp]),
        "mpirfft_memcpy_d2h": (i32, [vp, vp, sz, vp]),
        "mpirfft_stream_sync": (i32, [vp]),
        "mpirfft_set_pointwise_mode": (None, [i32]),
        "mpirfft_profile_enable": (None, [i32]),
        "mpirfft_profile_read": (i32, [C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(C.c_double), i32]),
        "mpirfft_measure_imad_rate": (C.c_double, [i32]),
        "mpirfft_launch_count": (u64, []),
        "mpirfft_launch_count_reset": (None, []),
        "new_mpn_mul": (None, [vp, vp, i64, vp, i64, u64, u64]),
        "mpirfft_mpn_mul": (None, [vp, vp, i64, vp, i64]),
        "new_mpn_mulmod_2expp1": (u64, [vp, vp, vp, u64, u64, vp]),
        "fft_mulmod_2expp1": (u64, [vp, vp, vp, i64, i64, vp]),
        "FFT_mulmod_2expp1": (None, [vp, vp, vp, i64, u64, u64]),
        "mpir_revbin": (u64, [u64, u64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L
