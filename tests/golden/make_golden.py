"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, i.e. the unmodified
/root/reference/mul_fft.c through oracle/shim; `new_mpn_mul` from the copy fixed at line 3246).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own (all its tests are randomised), so
these fixtures pin the CPU restatement (oracle/ssmul_oracle.c) and the CUDA path to outputs of the
reference itself.  Inputs are regenerated from seeds by the tests (tests/common.py); only outputs
(normalised residues as blocks, or product limbs / their SHA-256) are stored.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import loader as L                                    # noqa: E402
from common import rand_blocks, operand, ptr, cl, cul, block_to_int, int_to_block  # noqa: E402

TRANSFORMS = [
    # name, n, w, trunc, seed   (1-D, rr == ii, rs == 1)
    ("FFT_radix2", 16, 4, 0, 101), ("IFFT_radix2", 16, 4, 0, 102), ("FFT_radix2", 64, 3, 0, 103),
    ("FFT_radix2_truncate", 16, 8, 18, 104), ("FFT_radix2_truncate1", 16, 8, 6, 105),
    ("IFFT_radix2_truncate", 16, 8, 24, 106), ("IFFT_radix2_truncate1", 16, 8, 10, 107),
    ("FFT_radix2_truncate", 128, 1, 144, 108), ("IFFT_radix2_truncate", 128, 1, 130, 109),
    ("FFT_radix2_negacyclic", 32, 2, 0, 110), ("IFFT_radix2_negacyclic", 32, 2, 0, 111),
]
MFAS = [
    # inverse, n, w, n1, trunc, seed
    (0, 32, 2, 8, 0, 201), (1, 32, 2, 8, 0, 202), (0, 64, 1, 4, 72, 203), (1, 64, 1, 4, 72, 204),
    (0, 64, 2, 16, 96, 205), (1, 64, 2, 16, 96, 206), (0, 32, 128, 8, 48, 207), (1, 32, 128, 8, 48, 208),
]
SQRT2_MFAS = [
    # inverse, n, w, n1, trunc, seed   (FFT/IFFT_radix2_mfa_truncate_sqrt2: 4n blocks, trunc in (2n, 4n])
    (0, 64, 1, 8, 144, 301), (1, 64, 1, 8, 144, 302), (0, 64, 3, 8, 176, 303), (1, 64, 3, 8, 176, 304),
    (0, 64, 2, 8, 208, 305), (1, 64, 2, 8, 208, 306), (0, 256, 1, 16, 672, 307), (1, 256, 1, 16, 672, 308),
]
PRODUCTS6 = [
    # new_mpn_mul6: n1, n2, depth, w, kind
    (50, 41, 6, 1, "ones"), (3000, 2500, 8, 3, "uniform"), (3500, 4000, 6, 64, "ones"), (50000, 45000, 10, 3, "uniform"),
]
PRODUCTS = [
    # n1, n2, depth, w, kind      full limbs stored when small, sha256 otherwise
    (1, 1, 6, 1, "uniform"), (20, 13, 6, 1, "uniform"), (40, 40, 6, 2, "ones"), (700, 900, 7, 12, "runs"),
    (1500, 1500, 6, 64, "ones"), (6000, 6000, 6, 256, "uniform"), (1 << 16, 1 << 16, 12, 1, "uniform"),
    (1 << 16, 1 << 16, 12, 1, "ones"), (123457, 65521, 13, 1, "runs"), (100000, 100000, 12, 3, "uniform"),
]


def canon(blocks, l):
    return np.stack([int_to_block(block_to_int(b, l), l) for b in blocks])


def transform_inputs(name, n, w, trunc, seed):
    rng = np.random.default_rng(seed)
    l = n * w // 64
    data = rand_blocks(rng, 2 * n, l)
    if name == "FFT_radix2_truncate":
        data[trunc:] = 0
    return data, l


def mfa_inputs(inverse, n, w, n1, trunc, seed):
    rng = np.random.default_rng(seed)
    l = n * w // 64
    data = rand_blocks(rng, 2 * n, l)
    if trunc and not inverse:
        data[trunc:] = 0
    return data, l


def sqrt2_mfa_inputs(inverse, n, w, n1, trunc, seed):
    rng = np.random.default_rng(seed)
    l = n * w // 64
    data = rand_blocks(rng, 4 * n, l)
    if not inverse:
        data[trunc:] = 0
    return data, l


def sqrt2_valid(inverse, n, n1, trunc):
    """indices of the blocks the transform defines (mul_fft.c:2212-2355, 2593-2750)"""
    if inverse:
        return list(range(trunc))
    n2, trunc2 = 2 * n // n1, (trunc - 2 * n) // n1
    depth = n2.bit_length() - 1
    rows = list(range(n2)) + [n2 + int(format(s, "0%db" % depth)[::-1], 2) for s in range(trunc2)]
    return [r * n1 + c for r in rows for c in range(n1)]


def main():
    ref = L.load_ref(True)
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, n, w, trunc, seed in TRANSFORMS:
        data, l = transform_inputs(name, n, w, trunc, seed)
        s = L.Slab(2 * n, l, data)
        args = [s.ii, cl(1), s.ii, cl(n), cul(w), s.pt1, s.pt2, s.ptmp] + ([cl(trunc)] if trunc else [])
        getattr(ref, name)(*args)
        upto = trunc if trunc else 2 * n
        out["T_%s_%d_%d_%d" % (name, n, w, trunc)] = canon(s.all()[:upto], l)
    for inverse, n, w, n1, trunc, seed in MFAS:
        data, l = mfa_inputs(inverse, n, w, n1, trunc, seed)
        s = L.Slab(2 * n, l, data)
        name = ("I" if inverse else "") + "FFT_radix2_mfa" + ("_truncate" if trunc else "")
        getattr(ref, name)(*([s.ii, cl(n), cul(w), s.pt1, s.pt2, s.ptmp, cl(n1)] + ([cl(trunc)] if trunc else [])))
        out["M_%d_%d_%d_%d_%d" % (inverse, n, w, n1, trunc)] = canon(s.all(), l)
    for inverse, n, w, n1, trunc, seed in SQRT2_MFAS:
        data, l = sqrt2_mfa_inputs(inverse, n, w, n1, trunc, seed)
        s = L.Slab(4 * n, l, data)
        name = ("I" if inverse else "") + "FFT_radix2_mfa_truncate_sqrt2"
        getattr(ref, name)(s.ii, cl(n), cul(w), s.pt1, s.pt2, s.ptmp, cl(n1), cl(trunc))
        allb = s.all()
        out["S_%d_%d_%d_%d_%d" % (inverse, n, w, n1, trunc)] = canon([allb[k] for k in sqrt2_valid(inverse, n, n1, trunc)], l)
    for n1, n2, depth, w, kind in PRODUCTS6:
        a, b = operand(kind, n1, 0x5EED0001), operand(kind, n2, 0x5EED0002)
        r = np.zeros(n1 + n2, dtype=np.uint64)
        ref.new_mpn_mul6(ptr(r), ptr(a), cl(n1), ptr(b), cl(n2), cul(depth), cul(w))
        assert np.array_equal(r, L.gmp_mul(a, b)), "the reference's new_mpn_mul6 disagrees with mpn_mul?!"
        key = "P6_%d_%d_%d_%d_%s" % (n1, n2, depth, w, kind)
        if n1 + n2 <= 4096:
            out[key] = r
        else:
            out[key + "_sha256"] = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), dtype=np.uint8)
    for n1, n2, depth, w, kind in PRODUCTS:
        a, b = operand(kind, n1, 0x5EED0001), operand(kind, n2, 0x5EED0002)
        r = np.zeros(n1 + n2, dtype=np.uint64)
        ref.new_mpn_mul(ptr(r), ptr(a), cl(n1), ptr(b), cl(n2), cul(depth), cul(w))
        assert np.array_equal(r, L.gmp_mul(a, b)), "patched reference disagrees with mpn_mul?!"
        key = "P_%d_%d_%d_%d_%s" % (n1, n2, depth, w, kind)
        if n1 + n2 <= 4096:
            out[key] = r
        else:
            out[key + "_sha256"] = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote %d fixtures, %d bytes" % (len(out), os.path.getsize(os.path.join(HERE, "reference_outputs.npz"))))


if __name__ == "__main__":
    main()
