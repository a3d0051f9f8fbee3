"""oracle/loader.py -- TEST INFRASTRUCTURE ONLY.

ctypes access to the CPU checkers:
  * oracle/_ref/libmulfft_ref{,_fixed}.so : the UNMODIFIED reference translation unit
    (/root/reference/mul_fft.c) compiled against oracle/shim (see oracle/Makefile);
    `_fixed` carries the one-expression correction of mul_fft.c:3246.
  * oracle/libssmul_oracle.so             : the plain-C restatement (oracle/ssmul_oracle.c).
  * libgmp.so.10                          : GMP 6.3.0 runtime, used as the `mpn_mul` truth.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
The product package (mpir_fft_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u64p = C.POINTER(C.c_uint64)
u64pp = C.POINTER(u64p)


def _p(a):
    """numpy uint64 array -> ctypes pointer (array must stay alive)."""
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


_gmp = None


def gmp():
    global _gmp
    if _gmp is None:
        _gmp = C.CDLL("libgmp.so.10")
        _gmp.__gmpn_mul.restype = C.c_uint64
        _gmp.__gmpn_mul.argtypes = [u64p, u64p, C.c_long, u64p, C.c_long]
    return _gmp


def gmp_mul(a, b):
    """{a} * {b} -> len(a)+len(b) limbs, via GMP mpn_mul (the product is unique)."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    r = np.zeros(len(a) + len(b), dtype=np.uint64)
    if len(a) >= len(b):
        gmp().__gmpn_mul(_p(r), _p(a), len(a), _p(b), len(b))
    else:
        gmp().__gmpn_mul(_p(r), _p(b), len(b), _p(a), len(a))
    return r


def build(target="all"):
    subprocess.check_call(["make", "-s", "-C", HERE, target])


_libs = {}


def load_ref(fixed=True):
    """The compiled reference (None if oracle/_ref has not been built and cannot be)."""
    name = "libmulfft_ref_fixed.so" if fixed else "libmulfft_ref.so"
    if name not in _libs:
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            if os.path.exists("/root/reference/mul_fft.c"):
                build("ref")
            if not os.path.exists(path):
                _libs[name] = None
                return None
        _libs[name] = C.CDLL(path, mode=C.RTLD_LOCAL)
    return _libs[name]


def load_port():
    """The plain-C restatement oracle/ssmul_oracle.c (built on demand; gcc only)."""
    name = "libssmul_oracle.so"
    if name not in _libs:
        path = os.path.join(HERE, name)
        src = os.path.join(HERE, "ssmul_oracle.c")
        if (not os.path.exists(path)) or os.path.getmtime(path) < os.path.getmtime(src):
            build("port")
        _libs[name] = C.CDLL(path, mode=C.RTLD_LOCAL)
    return _libs[name]


class Slab:
    """A reference-style coefficient slab: `count` blocks of (l+1) limbs plus the
    pointer table `ii` the reference's FFT entry points take (mul_fft.c:3214-3221),
    plus the two scratch blocks t1, t2 and the unused `temp`."""

    def __init__(self, count, l, data=None):
        self.count, self.l, self.size = count, l, l + 1
        self.mem = np.zeros((count + 3) * self.size, dtype=np.uint64)
        if data is not None:
            self.mem[: count * self.size] = np.asarray(data, dtype=np.uint64).reshape(-1)
        base = self.mem.ctypes.data
        self.ii = (u64p * count)()
        for i in range(count):
            self.ii[i] = C.cast(base + 8 * i * self.size, u64p)
        self.t1 = u64p.from_address(0)
        self.t1 = C.cast(base + 8 * count * self.size, u64p)
        self.t2 = C.cast(base + 8 * (count + 1) * self.size, u64p)
        self.tmp = C.cast(base + 8 * (count + 2) * self.size, u64p)
        self.pt1, self.pt2, self.ptmp = u64p(), u64p(), u64p()
        self.pt1 = C.pointer(self.t1)
        self.pt2 = C.pointer(self.t2)
        self.ptmp = C.pointer(self.tmp)

    def get(self, i):
        """Block currently pointed to by ii[i] (the reference permutes pointers)."""
        addr = C.cast(self.ii[i], C.c_void_p).value
        off = (addr - self.mem.ctypes.data) // 8
        return self.mem[off: off + self.size].copy()

    def all(self):
        return np.stack([self.get(i) for i in range(self.count)])
