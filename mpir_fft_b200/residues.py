"""Residue checks for products too large to multiply a second time on the host.

An integer given as 64-bit limbs x_k (k counted from `first_index`) is reduced modulo several fixed
31-bit primes p:   sum_k x_k * W^k  (mod p),  W = 2^64 mod p.   For these primes the order of 2 is
large (W is not a small power of two with a short period, unlike 2^31-1 or 2^61-1, for which a
coefficient misplaced by a multiple of the period would go unnoticed), so a limb that lands at the
wrong position, in the wrong rank window, or with a wrong carry changes the residue.  a*b == r is
checked as  res(a) * res(b) == res(r)  (mod p) for every prime -- six primes, 186 bits.  The
residues of the result are additive over rank windows, so every rank reduces the limbs it owns and
one all-reduce (of 6 numbers) combines them; the tuple is also a position-dependent fingerprint
that does not depend on how the product was sharded (world-size independence).

Works on torch int64 tensors holding the limbs' bit patterns (device or host).
"""
import torch

# fixed 31-bit primes (deterministic: the fingerprints of different runs are comparable)
PRIMES = (2147483629, 2147483587, 2147483579, 2147483563, 2147483549, 2147483543)
_C = 1 << 12       # limbs per row of the two-level evaluation


def _powers(w, p, n):
    out, v = [], 1
    for _ in range(n):
        out.append(v)
        v = v * w % p
    return out


def _limbs_mod(x, p):
    """the unsigned 64-bit limbs (int64 bit patterns) reduced mod p (< 2^31), as int64 in [0, p)"""
    lo = x & 0xFFFFFFFF
    hi = (x >> 32) & 0xFFFFFFFF
    return ((hi * ((1 << 32) % p)) % p + lo) % p


def residues(limbs, first_index=0, chunk=1 << 24):
    """[sum_k limbs[k] * 2^(64 (first_index + k)) mod p for p in PRIMES] as python ints"""
    n = limbs.numel()
    dev = limbs.device
    out = []
    for p in PRIMES:
        W = pow(2, 64, p)
        pw = torch.tensor(_powers(W, p, _C), dtype=torch.int64, device=dev)
        WC = pow(W, _C, p)
        acc = 0
        for lo in range(0, n, chunk):
            x = limbs[lo:lo + chunk]
            m = x.numel()
            full = (m // _C) * _C
            part = 0
            if full:
                rows = (_limbs_mod(x[:full], p).view(-1, _C) * pw) % p          # products < 2^62
                rs = rows.sum(dim=1) % p                                         # 4096 terms < 2^31 each
                nr = rs.numel()
                # second level: sum_i rs[i] * (W^C)^i, again by rows of _C
                pad = (-nr) % _C
                if pad:
                    rs = torch.cat([rs, torch.zeros(pad, dtype=torch.int64, device=dev)])
                pw2 = torch.tensor(_powers(WC, p, _C), dtype=torch.int64, device=dev)
                r2 = ((rs.view(-1, _C) * pw2) % p).sum(dim=1) % p
                WCC = pow(WC, _C, p)
                v = 0
                for t in reversed(r2.cpu().tolist()):                            # Horner over <= chunk / 2^24 values
                    v = (v * WCC + t) % p
                part = v
            if m > full:
                tail = _limbs_mod(x[full:], p)
                tv = int(((tail * pw[:m - full]) % p).sum().item()) % p
                part = (part + tv * pow(W, full, p)) % p
            acc = (acc + part * pow(W, lo, p)) % p
        out.append(acc * pow(W, first_index, p) % p)
    return out


def combine(parts):
    """residues of a sum of disjoint windows from the windows' residues (lists of equal length)"""
    return [sum(col) % p for col, p in zip(zip(*parts), PRIMES)]


def product_matches(ra, rb, rr):
    return all((x * y - z) % p == 0 for x, y, z, p in zip(ra, rb, rr, PRIMES))
