"""CPU: the residue fingerprints used to check products that are too large to multiply twice
(mpir_fft_b200/residues.py) against Python big integers, and that they notice misplaced limbs."""
import numpy as np
import pytest
import torch

from mpir_fft_b200 import residues as R


def _int(x, first=0):
    return int.from_bytes(np.asarray(x, dtype=np.uint64).tobytes(), "little") << (64 * first)


@pytest.mark.parametrize("n,first", [(1, 0), (5, 0), (4096, 3), (4097, 0), (100000, 12345), (1 << 18, 7)])
def test_residues_match_bigint(n, first):
    rng = np.random.default_rng(n + first)
    x = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    got = R.residues(torch.from_numpy(x.view(np.int64)), first, chunk=1 << 16)
    v = _int(x, first)
    assert got == [v % p for p in R.PRIMES]


def test_product_check_and_window_additivity():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2 ** 64, 3000, dtype=np.uint64)
    b = rng.integers(0, 2 ** 64, 2500, dtype=np.uint64)
    prod = _int(a) * _int(b)
    r = np.frombuffer(prod.to_bytes(8 * 5500, "little"), dtype=np.uint64).copy()
    t = lambda v: torch.from_numpy(v.view(np.int64))      # noqa: E731
    ra, rb = R.residues(t(a)), R.residues(t(b))
    # the result reduced window by window, as the ranks of a sharded product do
    cuts = [0, 700, 701, 3000, 5500]
    parts = [R.residues(t(r[lo:hi].copy()), lo) for lo, hi in zip(cuts[:-1], cuts[1:])]
    rr = R.combine(parts)
    assert rr == R.residues(t(r)) and R.product_matches(ra, rb, rr)
    # a limb in the wrong place, two limbs swapped 61 positions apart (invisible to weights of period 61), a lost carry
    bad = r.copy(); bad[100], bad[161] = r[161], r[100]
    assert not R.product_matches(ra, rb, R.residues(t(bad)))
    bad = r.copy(); bad[2000] += np.uint64(1)
    assert not R.product_matches(ra, rb, R.residues(t(bad)))
    bad = np.roll(r, 1891)            # 31 * 61 limbs: invisible modulo 2^31-1 and 2^61-1
    assert not R.product_matches(ra, rb, R.residues(t(bad)))
