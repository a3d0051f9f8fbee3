"""Shared helpers for the test-suite (inputs, reference-style slabs, comparison)."""
import ctypes as C
import numpy as np

from oracle import loader as oracle
from schedsim import block_to_int, int_to_block   # noqa: F401

u64p = oracle.u64p
MASK = (1 << 64) - 1


def splitmix64(seed, n):
    """limb[k] = splitmix64(seed + k): counter-based uniform limbs."""
    z = (np.arange(n, dtype=np.uint64) + np.uint64(seed & MASK)) * np.uint64(0x9E3779B97F4A7C15)
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def operand(kind, n, seed=0x5EED0001):
    """synthetic operands: uniform, all-ones (worst-case carries), a single bit, long 0/1 runs"""
    if kind == "uniform":
        a = splitmix64(seed, n)
        a[-1] |= np.uint64(1)
        return a
    if kind == "ones":
        return np.full(n, MASK, dtype=np.uint64)
    if kind == "pow2":
        a = np.zeros(n, dtype=np.uint64)
        a[n - 1] = np.uint64(1) << np.uint64(seed % 64)
        return a
    if kind == "runs":                      # mpn_rrandom-style long runs of zeros and ones
        rng = np.random.default_rng(seed)
        bits = np.zeros(64 * n, dtype=np.uint8)
        pos, val = 0, 1
        while pos < 64 * n:
            ln = int(rng.integers(1, 4 * 64))
            bits[pos:pos + ln] = val
            pos += ln
            val ^= 1
        a = np.packbits(bits, bitorder="little").view(np.uint64).copy()
        a[-1] |= np.uint64(1)
        return a
    raise ValueError(kind)


def rand_blocks(rng, count, l, tops=True):
    """random (l+1)-limb blocks with small signed top limbs like the reference's rand_n
    (mul_fft.c:3770-3775)"""
    d = rng.integers(0, 2 ** 64, (count, l + 1), dtype=np.uint64)
    if tops:
        d[:, l] = rng.integers(-9, 10, count).astype(np.int64).view(np.uint64)
    else:
        d[:, l] = 0
    return d


def residues(blocks, l):
    return [block_to_int(b, l) for b in blocks]


def first_diff(a, b):
    bad = np.nonzero(np.asarray(a) != np.asarray(b))[0]
    return None if len(bad) == 0 else (int(bad[0]), int(bad[-1]), len(bad))


def ptr(a):
    return C.c_void_p(a.ctypes.data)


def cl(x):
    return C.c_long(int(x))


def cul(x):
    return C.c_ulong(int(x))
