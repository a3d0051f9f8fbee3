"""GPU parity tests: every call goes through the C ABI of libmpirfft_b200.so and is compared,
bit for bit, with the CPU checkers (GMP mpn_mul for products -- the product is unique -- and the
compiled reference oracle/_ref for the transform entry points, after mpn_normmod_2expp1 exactly as
the reference's own tests compare, mul_fft.c:4547-4557).  Integer work: the bar is bit-exact.
"""
import ctypes as C
import numpy as np
import pytest

import mpir_fft_b200 as M
from oracle import loader as oracle
from common import (operand, rand_blocks, residues, first_diff, ptr, cl, cul, block_to_int,
                    int_to_block, splitmix64)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = M.lib()
    assert L.mpirfft_device_count() > 0, "no CUDA device visible"
    M.init(0)
    return L


@pytest.fixture(scope="module")
def ref():
    r = oracle.load_ref(True)
    if r is None:
        pytest.skip("oracle/_ref not built")
    return r


def check_mul(n1, n2, depth, w, kind1="uniform", kind2="uniform"):
    a, b = operand(kind1, n1, 0x5EED0001), operand(kind2, n2, 0x5EED0002)
    got = M.new_mpn_mul(a, b, depth, w)
    want = oracle.gmp_mul(a, b)
    assert first_diff(got, want) is None, "new_mpn_mul %dx%d depth %d w %d (%s,%s): (first,last,count)=%s" % (
        n1, n2, depth, w, kind1, kind2, first_diff(got, want))


# ---- BASELINE.json configs 1-3 ----
def test_cfg1_2p16(lib):
    check_mul(1 << 16, 1 << 16, 12, 1)


def test_cfg1_matches_patched_reference(lib, ref):
    n = 1 << 16
    a, b = operand("uniform", n, 1), operand("uniform", n, 2)
    want = np.zeros(2 * n, dtype=np.uint64)
    ref.new_mpn_mul(ptr(want), ptr(a), cl(n), ptr(b), cl(n), cul(12), cul(1))
    assert first_diff(M.new_mpn_mul(a, b, 12, 1), want) is None


def test_cfg2_2p20(lib):
    check_mul(1 << 20, 1 << 20, 14, 1)


def test_cfg3_truncated_depth14_w2(lib):
    check_mul(3000000, 1700000, 14, 2)


def test_cfg3_truncated_depth15_w1(lib):
    check_mul(3000000, 1700000, 15, 1)


@pytest.mark.parametrize("kinds", [("ones", "ones"), ("runs", "runs"), ("pow2", "uniform"), ("ones", "uniform")])
def test_adversarial_operands_2p16(lib, kinds):
    check_mul(1 << 16, 1 << 16, 12, 1, *kinds)


def test_cfg2_all_ones(lib):
    check_mul(1 << 20, 1 << 20, 14, 1, "ones", "ones")


@pytest.mark.parametrize("case", [
    (1, 1, 6, 1), (20, 13, 6, 1), (40, 40, 6, 2), (200000, 3, 14, 1), (123457, 65521, 13, 1),
    (100000, 100000, 12, 3), (1 << 17, 1 << 17, 11, 8), (1 << 18, 1 << 18, 13, 1),
    (700, 900, 7, 12), (3000, 3000, 7, 96), (12000, 9000, 6, 512), (3 << 16, 3 << 16, 13, 2),
])
def test_odd_sizes(lib, case):
    check_mul(*case)


@pytest.mark.parametrize("case", [
    (30000, 28000, 6, 1024, "uniform"), (20000, 7, 6, 1024, "ones"), (5000, 6000, 5, 2048, "runs"),
    (1 << 22, 1 << 22, 16, 1, "uniform"),                  # 4M x 4M limbs in the ring of 1024 limbs
    (3000000, 2000001, 14, 8, "ones"),                     # ring of 2048 limbs, 2^14 x 2 coefficients
])
def test_big_ring_through_the_dropin_symbol(lib, case):
    """Rings above 512 limbs: new_mpn_mul runs the plan of the sharded multiplication on one rank
    (mul.c: exec_big)."""
    n1, n2, depth, w, kind = case
    check_mul(n1, n2, depth, w, kind, kind)


def test_size_independent_properties_cfg2(lib):
    """at full size: commutativity, and (a*b) mod small primes from the inputs alone"""
    n = 1 << 20
    a, b = operand("uniform", n, 11), operand("runs", n, 12)
    r1 = M.new_mpn_mul(a, b, 14, 1)
    r2 = M.new_mpn_mul(b, a, 14, 1)
    assert first_diff(r1, r2) is None
    for p in (2305843009213693951, 18446744073709551557, 4294967291):
        # fold limbs mod p with python ints in chunks
        def modp(x):
            v = int.from_bytes(x.tobytes(), "little")
            return v % p
        assert modp(r1) == (modp(a) * modp(b)) % p


def test_illegal_parameters_rejected():
    with pytest.raises(ValueError):
        M.new_mpn_mul(np.ones(1 << 17, np.uint64), np.ones(1 << 17, np.uint64), 11, 4)   # needs 4103 > 4096 coeffs
    with pytest.raises(ValueError):
        M.new_mpn_mul(np.ones(8, np.uint64), np.ones(8, np.uint64), 5, 1)               # 64 does not divide n*w


@pytest.mark.parametrize("n1,n2,kind", [(1 << 20, 1 << 20, "uniform"), (123457, 65521, "ones"), (3000000, 1700000, "runs"),
                                         (5000, 17, "uniform"), (1 << 16, 1 << 16, "uniform")])
def test_mpn_mul_wrapper_chooses_parameters(lib, n1, n2, kind):
    a, b = operand(kind, n1, 31), operand(kind, n2, 32)
    assert np.array_equal(M.mpn_mul(a, b), oracle.gmp_mul(a, b))


def test_squaring_uses_one_forward_transform(lib):
    n, depth, w = 1 << 20, 14, 1
    a = operand("uniform", n, 41)
    plan = M.MulPlan(n, n, depth, w)
    da, dr = lib.mpirfft_malloc_device(a.nbytes), lib.mpirfft_malloc_device(2 * a.nbytes)
    lib.mpirfft_memcpy_h2d(da, ptr(a), a.nbytes, None)
    lib.mpirfft_launch_count_reset()
    plan.exec_device(dr, da, da)
    r = np.zeros(2 * n, dtype=np.uint64)
    lib.mpirfft_memcpy_d2h(ptr(r), dr, r.nbytes, None)
    lib.mpirfft_stream_sync(None)
    assert lib.mpirfft_launch_count() < plan.launches
    lib.mpirfft_free_device(da); lib.mpirfft_free_device(dr)
    assert np.array_equal(r, oracle.gmp_mul(a, a))


# ---- mulmod 2^(64 l)+1 ----
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("l", [1, 3, 24, 64, 128, 256, 512])
def test_mulmod_vs_bigint(lib, l, mode):
    import random
    lib.mpirfft_set_pointwise_mode(mode)       # 0: default, 1: nested SS kernel, 2: Karatsuba blocks, 3: schoolbook blocks
    random.seed(l)
    NW = 64 * l
    p = (1 << NW) + 1
    A = [random.getrandbits(NW) for _ in range(20)] + [p - 2, p - 2, p - 1, p - 1, 0, 1, p - 2,
                                                       (1 << (NW // 2)) - 1, (1 << NW) - (1 << (NW // 2)), p - 2]
    B = [random.getrandbits(NW) for _ in range(20)] + [p - 2, 2, p - 1, 12345, 77, p - 2, 1 << (NW - 1),
                                                       (1 << (NW // 2)) - 1, (1 << NW) - (1 << (NW // 2)), random.getrandbits(NW)]
    a = np.stack([int_to_block(v, l) for v in A])
    b = np.stack([int_to_block(v, l) for v in B])
    out = M.mulmod_batch(a, b)
    lib.mpirfft_set_pointwise_mode(0)
    for k in range(len(A)):
        want = int_to_block(A[k] * B[k] % p, l)
        assert np.array_equal(out[k], want), "mulmod l=%d case %d mode %d" % (l, k, mode)


@pytest.mark.parametrize("l", [1024, 4096, 16384])
def test_mulmod_transform_path_vs_bigint(lib, l):
    """>= 250 limbs: batched negacyclic-transform path (mm.c; FFT_mulmod_2expp1, mul_fft.c:2998)"""
    import random
    random.seed(l)
    NW = 64 * l
    p = (1 << NW) + 1
    A = [random.getrandbits(NW) for _ in range(6)] + [p - 1, p - 2, 0, 1, (1 << NW) - 1, p - 1, (1 << (NW // 2)) - 1, p - 2]
    B = [random.getrandbits(NW) for _ in range(6)] + [p - 2, p - 2, 5, p - 1, (1 << NW) - 1, p - 1, (1 << (NW // 2)) - 1, 2]
    a = np.stack([int_to_block(v, l) for v in A])
    b = np.stack([int_to_block(v, l) for v in B])
    out = M.mulmod_batch(a, b)
    for k in range(len(A)):
        assert np.array_equal(out[k], int_to_block(A[k] * B[k] % p, l)), "mulmod l=%d case %d" % (l, k)


def test_cfg4_batch_vs_reference_fft_mulmod(lib, ref):
    """BASELINE configs[3] shape: independent 16 384-limb negacyclic products; a sample is compared
    with the compiled reference's fft_mulmod_2expp1 (outer n = 2^14, w = 64: bits = 2^20) and all of
    them with a*b mod p through GMP"""
    l, count = 16384, 48
    NW = 64 * l
    a = np.zeros((count, l + 1), np.uint64)
    b = np.zeros((count, l + 1), np.uint64)
    for k in range(count):
        a[k, :l] = operand("uniform" if k % 3 else "runs", l, 100 + k)
        b[k, :l] = operand("uniform" if k % 5 else "ones", l, 200 + k)
    out = M.mulmod_batch(a, b)
    assert not out[:, l].any()
    for k in range(count):
        full = oracle.gmp_mul(a[k, :l].copy(), b[k, :l].copy())
        lo = int.from_bytes(full[:l].tobytes(), "little")
        hi = int.from_bytes(full[l:].tobytes(), "little")
        want = (lo - hi) % ((1 << NW) + 1)
        assert int.from_bytes(out[k, :l].tobytes(), "little") == want, k
    for k in range(0, count, 16):
        r = np.zeros(l + 1, np.uint64)
        tt = np.zeros(2 * l + 2, np.uint64)
        ak, bk = a[k].copy(), b[k].copy()       # keep the copies alive across the call
        ref.fft_mulmod_2expp1(ptr(r), ptr(ak), ptr(bk), cl(1 << 14), cl(64), ptr(tt))
        assert np.array_equal(r[:l], out[k, :l]), k


def test_fft_mulmod_symbols_large(lib, ref):
    """the drop-in symbols on a 1024-limb residue: both go through the transform path"""
    l = 1024
    a = np.zeros(l + 1, np.uint64); b = np.zeros(l + 1, np.uint64)
    a[:l] = splitmix64(11, l); b[:l] = splitmix64(12, l)
    r1 = np.zeros(l + 1, np.uint64); r2 = np.zeros(l + 1, np.uint64); tt = np.zeros(2 * l + 2, np.uint64)
    lib.fft_mulmod_2expp1(ptr(r1), ptr(a), ptr(b), cl(1 << 10), cl(64), ptr(tt))
    a2, b2 = a.copy(), b.copy()
    ref.fft_mulmod_2expp1(ptr(r2), ptr(a2), ptr(b2), cl(1 << 10), cl(64), ptr(tt))
    assert np.array_equal(r1[:l], r2[:l])
    r3 = np.zeros(l + 1, np.uint64)
    lib.FFT_mulmod_2expp1(ptr(r3), ptr(a), ptr(b), cl(l), cul(4), cul(256))
    assert np.array_equal(r3[:l], r2[:l])


def test_product_with_nested_ss_pointwise(lib):
    lib.mpirfft_set_pointwise_mode(1)
    try:
        check_mul(1 << 20, 1 << 20, 14, 1)
        check_mul(1 << 16, 1 << 16, 12, 1, "ones", "ones")
    finally:
        lib.mpirfft_set_pointwise_mode(0)


def test_new_mpn_mulmod_2expp1_symbol(lib):
    l = 64
    a, b = splitmix64(5, l), splitmix64(6, l)
    r = np.zeros(l, dtype=np.uint64)
    tt = np.zeros(2 * l + 2, dtype=np.uint64)
    top = lib.new_mpn_mulmod_2expp1(ptr(r), ptr(a), ptr(b), 0, 64 * l, ptr(tt))
    p = (1 << (64 * l)) + 1
    want = int.from_bytes(a.tobytes(), "little") * int.from_bytes(b.tobytes(), "little") % p
    got = int.from_bytes(r.tobytes(), "little") + (int(top) << (64 * l))
    assert got == want
    # c = 1: i1 is 2^bits (limbs zero)
    z = np.zeros(l, dtype=np.uint64)
    top = lib.new_mpn_mulmod_2expp1(ptr(r), ptr(z), ptr(b), 1, 64 * l, ptr(tt))
    got = int.from_bytes(r.tobytes(), "little") + (int(top) << (64 * l))
    assert got == (p - 1) * int.from_bytes(b.tobytes(), "little") % p


# ---- transform entry points against the compiled reference (normalised, as its tests do) ----
def _run_pair(lib_fn, ref_fn, N, l, data, args_after_ii):
    s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
    ref_fn(*([s1.ii] + [x(s1) if callable(x) else x for x in args_after_ii]))
    lib_fn(*([s2.ii] + [x(s2) if callable(x) else x for x in args_after_ii]))
    return residues(s1.all(), l), residues(s2.all(), l)


T1 = lambda s: s.pt1   # noqa: E731
T2 = lambda s: s.pt2   # noqa: E731
TT = lambda s: s.ptmp  # noqa: E731
II = lambda s: s.ii    # noqa: E731


@pytest.mark.parametrize("n,w", [(16, 4), (64, 3), (32, 64), (8, 256)])
def test_fft_ifft_radix2(lib, ref, n, w):
    rng = np.random.default_rng(n * w)
    l, N = n * w // 64, 2 * n
    data = rand_blocks(rng, N, l)
    for name in ("FFT_radix2", "IFFT_radix2"):
        L_, R_ = getattr(lib, name), getattr(ref, name)
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        R_(s1.ii, cl(1), s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp)
        L_(s2.ii, cl(1), s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp)
        assert residues(s1.all(), l) == residues(s2.all(), l), name


@pytest.mark.parametrize("trunc", [2, 6, 16, 18, 24, 30, 32])
def test_truncated_1d(lib, ref, trunc):
    n, w = 16, 8
    rng = np.random.default_rng(trunc)
    l, N = n * w // 64, 2 * n
    for name, zero in (("FFT_radix2_truncate", True), ("FFT_radix2_truncate1", False),
                       ("IFFT_radix2_truncate", False), ("IFFT_radix2_truncate1", False)):
        data = rand_blocks(rng, N, l)
        if zero:
            data[trunc:] = 0
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        getattr(ref, name)(s1.ii, cl(1), s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, cl(trunc))
        getattr(lib, name)(s2.ii, cl(1), s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, cl(trunc))
        assert residues(s1.all()[:trunc], l) == residues(s2.all()[:trunc], l), name
    for name, zero in (("FFT_radix2_truncate_twiddle", True), ("FFT_radix2_truncate1_twiddle", False),
                       ("IFFT_radix2_truncate_twiddle", False), ("IFFT_radix2_truncate1_twiddle", False)):
        data = rand_blocks(rng, N, l)
        if zero:
            data[trunc:] = 0
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        a = (cl(1), cl(n), cul(w))
        getattr(ref, name)(s1.ii, *a, s1.pt1, s1.pt2, s1.ptmp, cl(1), cl(2), cl(3), cl(1), cl(trunc))
        getattr(lib, name)(s2.ii, *a, s2.pt1, s2.pt2, s2.ptmp, cl(1), cl(2), cl(3), cl(1), cl(trunc))
        assert residues(s1.all()[:trunc], l) == residues(s2.all()[:trunc], l), name


@pytest.mark.parametrize("n,w,n1,trunc", [(32, 2, 8, 0), (64, 1, 4, 0), (32, 2, 8, 48), (64, 1, 4, 72),
                                          (64, 2, 16, 96), (2048, 1, 64, 2176), (128, 32, 16, 160)])
def test_mfa(lib, ref, n, w, n1, trunc):
    rng = np.random.default_rng(n + trunc)
    l, N = n * w // 64, 2 * n
    n2 = N // n1
    depth = n2.bit_length() - 1
    rows = [int(lib.mpir_revbin(s, depth)) for s in range((trunc // n1) if trunc else n2)]
    for inverse in (0, 1):
        data = rand_blocks(rng, N, l)
        if trunc and not inverse:
            data[trunc:] = 0
        name = ("I" if inverse else "") + "FFT_radix2_mfa" + ("_truncate" if trunc else "")
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        extra = (cl(trunc),) if trunc else ()
        getattr(ref, name)(s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, cl(n1), *extra)
        getattr(lib, name)(s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, cl(n1), *extra)
        r1, r2 = residues(s1.all(), l), residues(s2.all(), l)
        if not inverse:
            idx = [i * n1 + j for i in rows for j in range(n1)]
        else:
            idx = list(range(trunc if trunc else N))
        assert [r1[k] for k in idx] == [r2[k] for k in idx], name


def test_mfa_roundtrip_scaled(lib):
    """IFFT(FFT(x)) == 2n * x (mul_fft.c:4938), at the cfg1 ring, truncated"""
    n, w, n1, trunc = 4096, 1, 64, 4224
    rng = np.random.default_rng(1)
    l, N = n * w // 64, 2 * n
    data = rand_blocks(rng, N, l, tops=False)
    data[trunc:] = 0
    s = oracle.Slab(N, l, data)
    lib.FFT_radix2_mfa_truncate(s.ii, cl(n), cul(w), s.pt1, s.pt2, s.ptmp, cl(n1), cl(trunc))
    lib.IFFT_radix2_mfa_truncate(s.ii, cl(n), cul(w), s.pt1, s.pt2, s.ptmp, cl(n1), cl(trunc))
    p = (1 << (64 * l)) + 1
    got = residues(s.all()[:trunc], l)
    want = [(block_to_int(data[k], l) * N) % p for k in range(trunc)]
    assert got == want


def test_butterflies_and_primitives(lib, ref):
    rng = np.random.default_rng(9)
    for (n, w) in [(64, 1), (64, 3), (256, 4), (64, 64)]:
        l = n * w // 64
        for i in (0, 1, 5, n - 1, n, n + 3, 2 * n - 1):
            blk = rand_blocks(rng, 2, l)
            for name in ("FFT_radix2_butterfly", "FFT_radix2_inverse_butterfly"):
                if name.endswith("inverse_butterfly") and i >= n:
                    continue
                outs = []
                for Lb in (ref, lib):
                    a, b = blk[0].copy(), blk[1].copy()
                    s, t = np.zeros(l + 1, np.uint64), np.zeros(l + 1, np.uint64)
                    getattr(Lb, name)(ptr(s), ptr(t), ptr(a), ptr(b), cl(i), cl(n), cul(w))
                    outs.append((block_to_int(s, l), block_to_int(t, l)))
                assert outs[0] == outs[1], (name, n, w, i)
            outs = []
            for Lb in (ref, lib):
                a = blk[0].copy()
                r = np.zeros(l + 1, np.uint64)
                Lb.FFT_twiddle(ptr(r), ptr(a), cl(i), cl(n), cul(w))
                outs.append(block_to_int(r, l))
            assert outs[0] == outs[1], ("FFT_twiddle", n, w, i)
        NW = n * w
        for (b1, b2) in [(0, 0), (1, 63), (64, 65), (NW - 1, NW + 1), (2 * NW - 1, 3 * NW + 7)]:
            for name in ("FFT_radix2_twiddle_butterfly", "FFT_radix2_twiddle_inverse_butterfly"):
                outs = []
                for Lb in (ref, lib):
                    a, b = blk[0].copy(), blk[1].copy()
                    s, t = np.zeros(l + 1, np.uint64), np.zeros(l + 1, np.uint64)
                    getattr(Lb, name)(ptr(s), ptr(t), ptr(a), ptr(b), cl(NW), cul(b1), cul(b2))
                    outs.append((block_to_int(s, l), block_to_int(t, l)))
                assert outs[0] == outs[1], (name, NW, b1, b2)
        for d in (0, 1, 31, 63):
            for name in ("mpn_mul_2expmod_2expp1", "mpn_div_2expmod_2expp1"):
                outs = []
                for Lb in (ref, lib):
                    a = blk[0].copy()
                    r = np.zeros(l + 1, np.uint64)
                    getattr(Lb, name)(ptr(r), ptr(a), cl(l), cul(d))
                    outs.append(block_to_int(r, l))
                assert outs[0] == outs[1], (name, l, d)
        for (x, y) in [(0, 0), (0, 1), (1, 0), (l // 2, l // 2), (l - 1, 1), (1, l - 1)]:
            if max(x, y) >= l:
                continue
            for name in ("mpn_lshB_sumdiffmod_2expp1", "mpn_sumdiff_rshBmod_2expp1"):
                outs = []
                for Lb in (ref, lib):
                    a, b = blk[0].copy(), blk[1].copy()
                    s, t = np.zeros(l + 1, np.uint64), np.zeros(l + 1, np.uint64)
                    getattr(Lb, name)(ptr(s), ptr(t), ptr(a), ptr(b), cl(l), cl(x), cl(y))
                    outs.append((block_to_int(s, l), block_to_int(t, l)))
                assert outs[0] == outs[1], (name, l, x, y)
        # normalise: limb-for-limb canonical form
        for top in (0, 1, -1, 5, -7):
            a = blk[0].copy()
            a[l] = np.int64(top).view(np.uint64)
            b = a.copy()
            ref.mpn_normmod_2expp1(ptr(a), cl(l))
            lib.mpn_normmod_2expp1(ptr(b), cl(l))
            assert np.array_equal(a, b)


def test_split_combine(lib, ref):
    rng = np.random.default_rng(4)
    for (total, bits, out) in [(100, 29, 1), (1000, 2045, 64), (777, 640, 12), (4096, 8185, 256), (50, 64, 3)]:
        limbs = rng.integers(0, 2 ** 64, total, dtype=np.uint64)
        length = (64 * total - 1) // bits + 1
        res = []
        for Lb in (ref, lib):
            s = oracle.Slab(length, out)
            Lb.FFT_split_bits.restype = C.c_long
            got = Lb.FFT_split_bits(s.ii, ptr(limbs), cl(total), cl(bits), cl(out))
            assert got == length
            res.append(s.all())
        assert np.array_equal(res[0], res[1]), ("split", total, bits, out)
        # combine back (coefficients are < 2^bits so the sum is the original integer)
        outs = []
        for Lb in (ref, lib):
            s = oracle.Slab(length, out, res[0])
            r = np.zeros(total, dtype=np.uint64)
            Lb.FFT_combine_bits(ptr(r), s.ii, cl(length), cl(bits), cl(out), cl(total))
            outs.append(r)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], limbs), ("combine", total, bits, out)
        # both ADD to what res holds on entry (mul_fft.c:185, 229-233); small addends, so that no carry leaves
        # one of the reference's per-coefficient windows
        outs = []
        for Lb in (ref, lib):
            s = oracle.Slab(length, out, res[0] >> np.uint64(1))
            r = (limbs >> np.uint64(2)).copy()
            Lb.FFT_combine_bits(ptr(r), s.ii, cl(length), cl(bits), cl(out), cl(total))
            outs.append(r)
        assert np.array_equal(outs[0], outs[1]), ("combine adds to res", total, bits, out)


# ---- the sqrt2 transforms and new_mpn_mul6 (mul_fft.c:591-700, 972, 2212, 2593, 3573) ----
@pytest.mark.parametrize("case", [
    # (n1, n2, depth, w, kinds): 2^(depth+1) < j1+j2-1 <= 2^(depth+2)
    (60000, 60000, 11, 1, ("uniform", "uniform")),          # l = 32: odd w, stage-per-launch path
    (60000, 60000, 10, 4, ("ones", "ones")),                # l = 64: fused tiles, even w
    (700000, 650000, 13, 1, ("uniform", "runs")),           # l = 128: odd w, two column classes on the fused path
    (3000000, 1700000, 14, 1, ("uniform", "uniform")),      # cfg3 sizes at l = 256 instead of 512
    (3000000, 1700000, 14, 1, ("ones", "ones")),
    (1 << 20, 1 << 20, 12, 6, ("uniform", "uniform")),      # l = 384, even w
    (1700000, 1700000, 13, 3, ("uniform", "ones")),         # l = 384, odd w
    (100000, 100000, 11, 3, ("runs", "uniform")), (123457, 65521, 11, 2, ("uniform", "ones")),
])
def test_new_mpn_mul6(lib, case):
    n1, n2, depth, w, (k1, k2) = case
    a, b = operand(k1, n1, 0x5EED0001), operand(k2, n2, 0x5EED0002)
    got = M.new_mpn_mul6(a, b, depth, w)
    want = oracle.gmp_mul(a, b)
    assert first_diff(got, want) is None, "new_mpn_mul6 %s: (first,last,count)=%s" % (case, first_diff(got, want))


def test_new_mpn_mul6_vs_golden(lib):
    """the committed new_mpn_mul6 fixtures (products of the compiled reference, tests/golden/make_golden.py)"""
    import hashlib, os
    from golden import make_golden as G
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
    for n1, n2, depth, w, kind in G.PRODUCTS6:
        a, b = operand(kind, n1, 0x5EED0001), operand(kind, n2, 0x5EED0002)
        r = M.new_mpn_mul6(a, b, depth, w)
        key = "P6_%d_%d_%d_%d_%s" % (n1, n2, depth, w, kind)
        if key in gold:
            assert np.array_equal(r, gold[key]), key
        else:
            assert hashlib.sha256(r.tobytes()).digest() == gold[key + "_sha256"].tobytes(), key


def test_new_mpn_mul6_matches_reference(lib, ref):
    """the reference's own big end-to-end test drives new_mpn_mul6 (mul_fft.c:5559)"""
    n1, n2, depth, w = 50000, 45000, 10, 3
    a, b = operand("uniform", n1, 1), operand("uniform", n2, 2)
    want = np.zeros(n1 + n2, dtype=np.uint64)
    ref.new_mpn_mul6(ptr(want), ptr(a), cl(n1), ptr(b), cl(n2), cul(depth), cul(w))
    assert first_diff(M.new_mpn_mul6(a, b, depth, w), want) is None
    assert first_diff(want, oracle.gmp_mul(a, b)) is None


@pytest.mark.parametrize("n,w,n1,trunc", [(64, 1, 8, 144), (64, 3, 8, 176), (64, 2, 8, 208), (256, 1, 16, 672),
                                          (2048, 2, 64, 4096 + 5 * 128), (4096, 1, 64, 8192 + 128), (8192, 1, 128, 4 * 8192)])
def test_mfa_sqrt2(lib, ref, n, w, n1, trunc):
    """FFT/IFFT_radix2_mfa_truncate_sqrt2 against the compiled reference (its tests: 4668, 4859)"""
    rng = np.random.default_rng(n + trunc + w)
    l, N = n * w // 64, 4 * n
    n2, trunc2 = 2 * n // n1, (trunc - 2 * n) // n1
    depth = n2.bit_length() - 1
    rows = list(range(n2)) + [n2 + int(lib.mpir_revbin(s, depth)) for s in range(trunc2)]
    for inverse in (0, 1):
        data = rand_blocks(rng, N, l)
        if not inverse:
            data[trunc:] = 0
        full = (trunc == 4 * n)         # the untruncated entry points (2078, 2461) take no trunc
        name = ("I" if inverse else "") + ("FFT_radix2_mfa_sqrt2" if full else "FFT_radix2_mfa_truncate_sqrt2")
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        extra = () if full else (cl(trunc),)
        getattr(ref, name)(s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, cl(n1), *extra)
        getattr(lib, name)(s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, cl(n1), *extra)
        r1, r2 = residues(s1.all(), l), residues(s2.all(), l)
        idx = [i * n1 + j for i in rows for j in range(n1)] if not inverse else list(range(trunc))
        assert [r1[k] for k in idx] == [r2[k] for k in idx], name


def test_sqrt2_butterflies(lib, ref):
    rng = np.random.default_rng(19)
    for (n, w) in [(256, 1), (256, 3), (64, 4 + 1), (1024, 1)]:
        if (n * w) % 256:
            continue
        l = n * w // 64
        for i in (1, 3, 5, n - 1, n + 1, 2 * n - 1, 2 * n + 1, 4 * n - 1):
            blk = rand_blocks(rng, 2, l)
            for name in ("FFT_radix2_butterfly_sqrt2", "FFT_radix2_inverse_butterfly_sqrt2"):
                if name.startswith("FFT_radix2_inverse") and i >= 2 * n:
                    continue                                   # the reference's exponent goes negative there (653)
                outs = []
                for Lb in (ref, lib):
                    a, b = blk[0].copy(), blk[1].copy()
                    s, t, tmp = np.zeros(l + 1, np.uint64), np.zeros(l + 1, np.uint64), np.zeros(2 * l + 2, np.uint64)
                    getattr(Lb, name)(ptr(s), ptr(t), ptr(a), ptr(b), cl(i), cl(n), cul(w), ptr(tmp))
                    outs.append((block_to_int(s, l), block_to_int(t, l)))
                assert outs[0] == outs[1], (name, n, w, i)
            outs = []
            for Lb in (ref, lib):
                a = blk[0].copy()
                r, tmp = np.zeros(l + 1, np.uint64), np.zeros(2 * l + 2, np.uint64)
                Lb.FFT_twiddle_sqrt2(ptr(r), ptr(a), cl(i), cl(n), cul(w), ptr(tmp))
                outs.append(block_to_int(r, l))
            assert outs[0] == outs[1], ("FFT_twiddle_sqrt2", n, w, i)


# ---- direct-symbol checks of the entry points that were only reached through other paths ----
@pytest.mark.parametrize("n,w", [(16, 8), (64, 2), (32, 64), (8, 256), (64, 1), (64, 3), (4096, 1), (1024, 5)])
def test_negacyclic_symbols(lib, ref, n, w):
    """FFT/IFFT_radix2_negacyclic (mul_fft.c:1290, 1861) against the compiled reference; odd w twists by
    powers of sqrt2^w (the reference's test_fft_ifft_negacyclic runs depth 11, w 1)"""
    rng = np.random.default_rng(n * w + 1)
    l, N = n * w // 64, 2 * n
    for name in ("FFT_radix2_negacyclic", "IFFT_radix2_negacyclic"):
        data = rand_blocks(rng, N, l)
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        getattr(ref, name)(s1.ii, cl(1), s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp)
        getattr(lib, name)(s2.ii, cl(1), s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp)
        assert residues(s1.all(), l) == residues(s2.all(), l), name


@pytest.mark.parametrize("n,w,is_,twiddle", [(16, 8, 1, (1, 2, 3, 1)), (32, 4, 2, (2, 3, 5, 2)), (16, 16, 4, (4, 0, 7, 1))])
def test_twiddle_symbols(lib, ref, n, w, is_, twiddle):
    """FFT/IFFT_radix2_twiddle (mul_fft.c:1397, 1964; mul_fft.h:73-79), strided, with the z^(r c) twist"""
    rng = np.random.default_rng(n + w + is_)
    l, N = n * w // 64, 2 * n
    ws, r, c, rs = twiddle
    for name in ("FFT_radix2_twiddle", "IFFT_radix2_twiddle"):
        data = rand_blocks(rng, N * is_, l)
        s1, s2 = oracle.Slab(N * is_, l, data), oracle.Slab(N * is_, l, data)
        getattr(ref, name)(s1.ii, cl(is_), cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, cl(ws), cl(r), cl(c), cl(rs))
        getattr(lib, name)(s2.ii, cl(is_), cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, cl(ws), cl(r), cl(c), cl(rs))
        r1, r2 = residues(s1.all(), l), residues(s2.all(), l)
        assert [r1[k * is_] for k in range(N)] == [r2[k * is_] for k in range(N)], name
        others = [k for k in range(N * is_) if k % is_]
        assert [r2[k] for k in others] == [block_to_int(data[k], l) for k in others], name + ": untouched blocks"


@pytest.mark.parametrize("n,w,n1,trunc", [(16384, 1, 128, 16640), (16384, 2, 128, 18432), (32768, 1, 128, 18432)])
def test_mfa_at_bench_rings(lib, ref, n, w, n1, trunc):
    """the fused tile passes at the cfg2 / cfg3 rings (l = 256, 512), pinned below product level:
    FFT/IFFT_radix2_mfa_truncate against the compiled reference, limb for limb after normalisation"""
    rng = np.random.default_rng(n + trunc)
    l, N = n * w // 64, 2 * n
    n2 = N // n1
    depth = n2.bit_length() - 1
    rows = [int(lib.mpir_revbin(s, depth)) for s in range(trunc // n1)]
    for inverse in (0, 1):
        data = rand_blocks(rng, N, l)
        if not inverse:
            data[trunc:] = 0
        name = ("I" if inverse else "") + "FFT_radix2_mfa_truncate"
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        getattr(ref, name)(s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, cl(n1), cl(trunc))
        getattr(lib, name)(s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, cl(n1), cl(trunc))
        idx = [i * n1 + j for i in rows for j in range(n1)] if not inverse else list(range(trunc))
        a1, a2 = s1.all(), s2.all()
        for k in idx[::97] + idx[-3:]:
            assert block_to_int(a1[k], l) == block_to_int(a2[k], l), (name, k)
        sample = idx if len(idx) < 4000 else idx[::7]
        assert [block_to_int(a1[k], l) for k in sample] == [block_to_int(a2[k], l) for k in sample], name


@pytest.mark.parametrize("n,w,trunc", [(64, 1, 256), (64, 1, 130), (64, 3, 192), (16, 4, 50), (1024, 1, 4096), (2048, 1, 2 * 2048 + 1234)])
def test_sqrt2_1d_symbols(lib, ref, n, w, trunc):
    """FFT/IFFT_radix2_sqrt2, FFT/IFFT_radix2_truncate_sqrt2 (mul_fft.c:839, 1488, 1230, 1792) vs the compiled reference"""
    rng = np.random.default_rng(n + w + trunc)
    l, N = n * w // 64, 4 * n
    full = trunc == 4 * n
    for inverse in (0, 1):
        data = rand_blocks(rng, N, l)
        if not inverse:
            data[trunc:] = 0
        name = ("I" if inverse else "") + ("FFT_radix2_sqrt2" if full else "FFT_radix2_truncate_sqrt2")
        s1, s2 = oracle.Slab(N, l, data), oracle.Slab(N, l, data)
        extra = () if full else (cl(trunc),)
        getattr(ref, name)(s1.ii, cl(1), s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, *extra)
        getattr(lib, name)(s2.ii, cl(1), s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, *extra)
        assert residues(s1.all()[:trunc], l) == residues(s2.all()[:trunc], l), name
