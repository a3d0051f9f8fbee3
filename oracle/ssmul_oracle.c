/* oracle/ssmul_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not product code, never linked into or
 * loaded by libmpirfft_b200.so; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may use it.
 *
 * A plain-C (no GMP, no MPIR) CPU restatement of the Schoenhage-Strassen path of
 * wbhart/mpir-fft (/root/reference/mul_fft.c), written in a deliberately different style from
 * the reference: every coefficient is kept as a CANONICAL residue mod p = 2^(64 l)+1 (l limbs
 * plus a top limb that is 0, or 1 with a zero body) and all arithmetic goes through three
 * primitives (add, sub, multiply by 2^e).  The recursions follow the reference routine named in
 * each comment, so that outputs agree with the reference after mpn_normmod_2expp1 -- which is how
 * the reference's own tests compare (mul_fft.c:4547-4557, 5089-5091).
 *
 * PARITY PINNING: the reference ships no golden vectors (all its tests are randomised).  This
 * restatement is pinned (tests/test_oracle.py) against (a) the UNMODIFIED reference compiled into
 * oracle/_ref through oracle/shim, function by function on seeded inputs, (b) GMP 6.3 mpn_mul for
 * whole products, and (c) the fixtures under tests/golden/ generated from (a).
 */
#include "ssmul_oracle.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------- limb vectors ------------------------------------------- */
static limb add_n(limb *r, const limb *a, const limb *b, long n)
{ limb c = 0; for (long i = 0; i < n; i++) { u128 t = (u128) a[i] + b[i] + c; r[i] = (limb) t; c = (limb)(t >> 64); } return c; }
static limb sub_n(limb *r, const limb *a, const limb *b, long n)
{ limb c = 0; for (long i = 0; i < n; i++) { u128 t = (u128) a[i] - b[i] - c; r[i] = (limb) t; c = (limb)(t >> 64) & 1; } return c; }
static limb add_1(limb *r, long n, limb v)
{ for (long i = 0; i < n && v; i++) { limb t = r[i] + v; v = t < v; r[i] = t; } return v; }
static limb sub_1(limb *r, long n, limb v)
{ for (long i = 0; i < n && v; i++) { limb t = r[i]; r[i] = t - v; v = t < v; } return v; }
static int is_zero(const limb *a, long n) { for (long i = 0; i < n; i++) if (a[i]) return 0; return 1; }

/* ------------------------------- residues mod 2^(64 l)+1 -------------------------------- */
/* mpn_normmod_2expp1 (mul_fft.c:272-294): any (l+1)-limb two's-complement block -> canonical */
void orc_normalise(limb *t, long l)
{
   for (int it = 0; it < 4; it++)
   {
      slimb hi = (slimb) t[l];
      if (hi == 0) return;
      if (hi == 1 && is_zero(t, l)) return;
      t[l] = 0;
      if (hi > 0) { if (sub_1(t, l, (limb) hi)) t[l] = (limb)(slimb) -1; }
      else        { if (add_1(t, l, (limb)(-hi))) t[l] = 1; }
   }
}

/* r = a + b (canonical in, canonical out) */
static void mod_add(limb *r, const limb *a, const limb *b, long l)
{
   limb c = add_n(r, a, b, l);
   r[l] = a[l] + b[l] + c;
   orc_normalise(r, l);
}
/* r = a - b */
static void mod_sub(limb *r, const limb *a, const limb *b, long l)
{
   limb c = sub_n(r, a, b, l);
   r[l] = a[l] - b[l] - c;
   orc_normalise(r, l);
}
/* r = a * 2^e, e taken mod 2*64*l; r must not alias a.  2^(64 l) == -1, so the product is a
 * negacyclic bit rotation (the fused shift of mul_fft.c:303-385, 470-488, 926-957). */
void orc_mul_2exp(limb *r, const limb *a, long l, unsigned long e)
{
   unsigned long NW = 64ul*(unsigned long) l;
   int neg = 0; long y; unsigned bs;
   limb *lo = (limb *) calloc((size_t) l + 1, sizeof(limb)), *hi = (limb *) calloc((size_t) l + 1, sizeof(limb));
   e %= 2*NW; if (e >= NW) { e -= NW; neg = 1; }
   y = (long)(e/64); bs = (unsigned)(e % 64);
   /* value = a_body + a_top*2^NW == a_body - a_top.  Shift the (l+1)-limb integer a_body left by
      e bits inside a 2l-limb window: low l limbs -> lo, upper part -> hi, result lo - hi. */
   for (long k = 0; k < l; k++)
   {
      limb v = a[k], w0 = bs ? (v << bs) : v, w1 = bs ? (v >> (64 - bs)) : 0;
      long p0 = k + y, p1 = k + y + 1;
      if (p0 < l) lo[p0] |= w0; else hi[p0 - l] |= w0;
      if (w1) { if (p1 < l) lo[p1] |= w1; else hi[p1 - l] |= w1; }
   }
   if (sub_n(r, lo, hi, l)) r[l] = (limb)(slimb) -1; else r[l] = 0;
   orc_normalise(r, l);
   if (a[l])
   {  /* canonical top == 1 means a == 2^NW == -1: subtract 2^e */
      limb *one = (limb *) calloc((size_t) l + 1, sizeof(limb));
      one[y] = (limb) 1 << bs;
      mod_sub(lo, r, one, l); memcpy(r, lo, sizeof(limb)*((size_t) l + 1));
      free(one);
   }
   if (neg)
   {
      memset(hi, 0, sizeof(limb)*((size_t) l + 1));
      mod_sub(lo, hi, r, l); memcpy(r, lo, sizeof(limb)*((size_t) l + 1));
   }
   free(lo); free(hi);
}

/* --------------------------------- butterflies ------------------------------------------ */
typedef struct { long l; unsigned long NW; limb *t0, *t1, *t2; } ring;

static void ring_init(ring *R, long l)
{
   R->l = l; R->NW = 64ul*(unsigned long) l;
   R->t0 = (limb *) malloc(sizeof(limb)*((size_t) l + 1));
   R->t1 = (limb *) malloc(sizeof(limb)*((size_t) l + 1));
   R->t2 = (limb *) malloc(sizeof(limb)*((size_t) l + 1));
}
static void ring_clear(ring *R) { free(R->t0); free(R->t1); free(R->t2); }
static void cpy(ring *R, limb *d, const limb *s) { memcpy(d, s, sizeof(limb)*((size_t) R->l + 1)); }

/* [a,b] -> [a+b, 2^e (a-b)]   FFT_radix2_butterfly, mul_fft.c:553-576 */
static void bfly_fwd(ring *R, limb *a, limb *b, unsigned long e)
{
   mod_add(R->t0, a, b, R->l); mod_sub(R->t1, a, b, R->l);
   cpy(R, a, R->t0); orc_mul_2exp(b, R->t1, R->l, e);
}
/* [a,b] -> [a + 2^-e b, a - 2^-e b]   FFT_radix2_inverse_butterfly, mul_fft.c:639-652 */
static void bfly_inv(ring *R, limb *a, limb *b, unsigned long e)
{
   orc_mul_2exp(R->t2, b, R->l, 2*R->NW - (e % (2*R->NW)));
   mod_add(R->t0, a, R->t2, R->l); mod_sub(R->t1, a, R->t2, R->l);
   cpy(R, a, R->t0); cpy(R, b, R->t1);
}
/* [s,t] -> [2^b1 (s+t), 2^b2 (s-t)]   FFT_radix2_twiddle_butterfly, mul_fft.c:517-548 */
static void bfly_twiddle(ring *R, limb *s, limb *t, unsigned long b1, unsigned long b2)
{
   mod_add(R->t0, s, t, R->l); mod_sub(R->t1, s, t, R->l);
   orc_mul_2exp(s, R->t0, R->l, b1); orc_mul_2exp(t, R->t1, R->l, b2);
}
/* [a,b] -> [2^-b1 a + 2^-b2 b, 2^-b1 a - 2^-b2 b]   mul_fft.c:721-752 */
static void bfly_twiddle_inv(ring *R, limb *a, limb *b, unsigned long b1, unsigned long b2)
{
   unsigned long M2 = 2*R->NW;
   orc_mul_2exp(R->t0, a, R->l, M2 - b1 % M2); orc_mul_2exp(R->t1, b, R->l, M2 - b2 % M2);
   mod_add(a, R->t0, R->t1, R->l); mod_sub(b, R->t0, R->t1, R->l);
}

/* ------------------------- 1-D transforms on block arrays -------------------------------- */
/* ii: array of block pointers, stride `is` (in pointers); twist 2^(ws*c*(r + rs*f)); ws == 0
 * gives the untwisted routines.  All in place (the reference's pointer swaps are not modelled:
 * only values are pinned). */
typedef struct { ring R; long is; unsigned long ws, c; } tctx;
#define B(k) (ii[(k)*T->is])

/* FFT_radix2 (786-827) / FFT_radix2_twiddle (1397-1442) */
static void fft_full(tctx *T, limb **ii, long n, unsigned long w, unsigned long r, unsigned long rs)
{
   if (n == 1) { bfly_twiddle(&T->R, B(0), B(1), r*T->c*T->ws, (r + rs)*T->c*T->ws); return; }
   for (long i = 0; i < n; i++) bfly_fwd(&T->R, B(i), B(n + i), (unsigned long) i*w);
   fft_full(T, ii, n/2, 2*w, r, 2*rs);
   fft_full(T, ii + n*T->is, n/2, 2*w, r + rs, 2*rs);
}
/* FFT_radix2_truncate1 (1028-1074) / _twiddle (1076-1122) */
static void fft_trunc1(tctx *T, limb **ii, long n, unsigned long w, unsigned long r, unsigned long rs, long trunc)
{
   if (trunc == 2*n) { fft_full(T, ii, n, w, r, rs); return; }
   if (trunc <= n)
   {
      for (long i = 0; i < n; i++) { mod_add(T->R.t0, B(i), B(i + n), T->R.l); cpy(&T->R, B(i), T->R.t0); }
      fft_trunc1(T, ii, n/2, 2*w, r, 2*rs, trunc);
   } else
   {
      for (long i = 0; i < n; i++) bfly_fwd(&T->R, B(i), B(n + i), (unsigned long) i*w);
      fft_full(T, ii, n/2, 2*w, r, 2*rs);
      fft_trunc1(T, ii + n*T->is, n/2, 2*w, r + rs, 2*rs, trunc - n);
   }
}
/* FFT_radix2_truncate (1128-1177) / _twiddle (1179-1228): inputs past trunc are zero */
static void fft_trunc(tctx *T, limb **ii, long n, unsigned long w, unsigned long r, unsigned long rs, long trunc)
{
   if (trunc == 2*n) { fft_full(T, ii, n, w, r, rs); return; }
   if (trunc <= n) { fft_trunc(T, ii, n/2, 2*w, r, 2*rs, trunc); return; }
   for (long i = 0; i < trunc - n; i++) bfly_fwd(&T->R, B(i), B(n + i), (unsigned long) i*w);
   for (long i = trunc; i < 2*n; i++) orc_mul_2exp(B(i), B(i - n), T->R.l, (unsigned long)(i - n)*w);
   fft_full(T, ii, n/2, 2*w, r, 2*rs);
   fft_trunc1(T, ii + n*T->is, n/2, 2*w, r + rs, 2*rs, trunc - n);
}
/* IFFT_radix2 (1444-1486) / IFFT_radix2_twiddle (1964-2010) */
static void ifft_full(tctx *T, limb **ii, long n, unsigned long w, unsigned long r, unsigned long rs)
{
   if (n == 1) { bfly_twiddle_inv(&T->R, B(0), B(1), r*T->c*T->ws, (r + rs)*T->c*T->ws); return; }
   ifft_full(T, ii, n/2, 2*w, r, 2*rs);
   ifft_full(T, ii + n*T->is, n/2, 2*w, r + rs, 2*rs);
   for (long i = 0; i < n; i++) bfly_inv(&T->R, B(i), B(n + i), (unsigned long) i*w);
}
/* IFFT_radix2_truncate1 (1538-1602) / _twiddle (1604-1668) */
static void ifft_trunc1(tctx *T, limb **ii, long n, unsigned long w, unsigned long r, unsigned long rs, long trunc)
{
   ring *R = &T->R; unsigned long M2 = 2*R->NW;
   if (trunc == 2*n) { ifft_full(T, ii, n, w, r, rs); return; }
   if (trunc <= n)
   {
      for (long i = trunc; i < n; i++)
      {  /* (ii[i] + ii[i+n]) / 2 */
         mod_add(R->t0, B(i), B(i + n), R->l); orc_mul_2exp(B(i), R->t0, R->l, M2 - 1);
      }
      ifft_trunc1(T, ii, n/2, 2*w, r, 2*rs, trunc);
      for (long i = 0; i < trunc; i++)
      {  /* 2 ii[i] - ii[n+i]  (mpn_addsub_n, 1565/1631) */
         mod_add(R->t0, B(i), B(i), R->l); mod_sub(R->t1, R->t0, B(n + i), R->l); cpy(R, B(i), R->t1);
      }
      return;
   }
   ifft_full(T, ii, n/2, 2*w, r, 2*rs);
   for (long i = trunc - n; i < n; i++)
   {  /* d = a - b; b' = 2^{iw} d; a' = a + d   (1575-1580) */
      mod_sub(R->t0, B(i), B(i + n), R->l);
      mod_add(R->t1, B(i), R->t0, R->l); cpy(R, B(i), R->t1);
      orc_mul_2exp(B(i + n), R->t0, R->l, (unsigned long) i*w);
   }
   ifft_trunc1(T, ii + n*T->is, n/2, 2*w, r + rs, 2*rs, trunc - n);
   for (long i = 0; i < trunc - n; i++) bfly_inv(R, B(i), B(n + i), (unsigned long) i*w);
}
/* IFFT_radix2_truncate (1674-1731) / _twiddle (1733-1790) */
static void ifft_trunc(tctx *T, limb **ii, long n, unsigned long w, unsigned long r, unsigned long rs, long trunc)
{
   ring *R = &T->R;
   if (trunc == 2*n) { ifft_full(T, ii, n, w, r, rs); return; }
   if (trunc <= n)
   {
      ifft_trunc(T, ii, n/2, 2*w, r, 2*rs, trunc);
      for (long i = 0; i < trunc; i++) { mod_add(R->t0, B(i), B(i), R->l); cpy(R, B(i), R->t0); }
      return;
   }
   ifft_full(T, ii, n/2, 2*w, r, 2*rs);
   for (long i = trunc; i < 2*n; i++) orc_mul_2exp(B(i), B(i - n), R->l, (unsigned long)(i - n)*w);
   ifft_trunc1(T, ii + n*T->is, n/2, 2*w, r + rs, 2*rs, trunc - n);
   for (long i = 0; i < trunc - n; i++) bfly_inv(R, B(i), B(n + i), (unsigned long) i*w);
   for (long i = trunc - n; i < n; i++) { mod_add(R->t0, B(i), B(i), R->l); cpy(R, B(i), R->t0); }
}
#undef B

static unsigned long revbin(unsigned long in, unsigned bits)     /* mpir_revbin, 63-79 */
{ unsigned long o = 0; for (unsigned i = 0; i < bits; i++) { o = (o << 1) | (in & 1); in >>= 1; } return o; }
static unsigned ilog2(unsigned long x) { unsigned b = 0; while ((1ul << b) < x) b++; return b; }

static void permute_revbin(limb **ii, long is, long count, unsigned bits)
{
   for (long j = 0; j < count; j++)
   {
      long s = (long) revbin((unsigned long) j, bits);
      if (j < s) { limb *t = ii[j*is]; ii[j*is] = ii[s*is]; ii[s*is] = t; }
   }
}

/* kind: 0 FFT 1 FFT_truncate 2 FFT_truncate1 3 IFFT 4 IFFT_truncate 5 IFFT_truncate1; blocks are
 * normalised first and left canonical.  ii[k*is], k < 2n. */
void orc_transform(int kind, limb **ii, long is, long n, unsigned long w, unsigned long ws,
                   unsigned long r, unsigned long c, unsigned long rs, long trunc)
{
   tctx T; long l = (long)((unsigned long) n*w/64);
   ring_init(&T.R, l); T.is = is; T.ws = ws; T.c = c;
   for (long k = 0; k < 2*n; k++) orc_normalise(ii[k*is], l);
   switch (kind)
   {
   case 0: fft_full(&T, ii, n, w, r, rs); break;
   case 1: fft_trunc(&T, ii, n, w, r, rs, trunc); break;
   case 2: fft_trunc1(&T, ii, n, w, r, rs, trunc); break;
   case 3: ifft_full(&T, ii, n, w, r, rs); break;
   case 4: ifft_trunc(&T, ii, n, w, r, rs, trunc); break;
   case 5: ifft_trunc1(&T, ii, n, w, r, rs, trunc); break;
   }
   ring_clear(&T.R);
}

/* FFT_radix2_mfa (2021-2068), FFT_radix2_mfa_truncate (2357-2409; trunc != 0), and the inverse
 * twins IFFT_radix2_mfa (2411-2459), IFFT_radix2_mfa_truncate (2925-2979).  The pointer table is
 * permuted exactly as the reference does, so ii[k] afterwards addresses result k. */
void orc_mfa(int inverse, limb **ii, long n, unsigned long w, long n1, long trunc)
{
   long n2 = 2*n/n1, l = (long)((unsigned long) n*w/64), rows = trunc ? trunc/n1 : n2;
   unsigned d1 = ilog2((unsigned long) n2), d2 = ilog2((unsigned long) n1);
   tctx T; ring_init(&T.R, l);
   for (long k = 0; k < 2*n; k++) orc_normalise(ii[k], l);
   if (!inverse)
   {
      for (long i = 0; i < n1; i++)
      {
         T.is = n1; T.ws = w; T.c = (unsigned long) i;
         if (trunc) fft_trunc(&T, ii + i, n2/2, w*(unsigned long) n1, 0, 1, rows);
         else fft_full(&T, ii + i, n2/2, w*(unsigned long) n1, 0, 1);
         permute_revbin(ii + i, n1, n2, d1);
      }
      for (long s = 0; s < rows; s++)
      {
         long i = (long) revbin((unsigned long) s, d1);
         T.is = 1; T.ws = 0; T.c = 0;
         fft_full(&T, ii + i*n1, n1/2, w*(unsigned long) n2, 0, 0);
         permute_revbin(ii + i*n1, 1, n1, d2);
      }
   } else
   {
      for (long s = 0; s < rows; s++)
      {
         long i = (long) revbin((unsigned long) s, d1);
         T.is = 1; T.ws = 0; T.c = 0;
         permute_revbin(ii + i*n1, 1, n1, d2);
         ifft_full(&T, ii + i*n1, n1/2, w*(unsigned long) n2, 0, 0);
      }
      for (long i = 0; i < n1; i++)
      {
         T.is = n1; T.ws = w; T.c = (unsigned long) i;
         permute_revbin(ii + i, n1, n2, d1);
         if (trunc) ifft_trunc(&T, ii + i, n2/2, w*(unsigned long) n1, 0, 1, rows);
         else ifft_full(&T, ii + i, n2/2, w*(unsigned long) n1, 0, 1);
      }
   }
   ring_clear(&T.R);
}

/* ------------------------------ split / combine ------------------------------------------ */
static limb bits_at(const limb *src, long n, unsigned long off)
{
   unsigned long q = off/64; unsigned r = (unsigned)(off % 64); limb v;
   if ((long) q >= n) return 0;
   v = src[q] >> r;
   if (r && (long) q + 1 < n) v |= src[q + 1] << (64 - r);
   return v;
}
/* FFT_split_bits (115-170): coefficient i = bits [i*bits, (i+1)*bits) */
long orc_split_bits(limb **poly, const limb *limbs, long total, long bits, long out)
{
   long length = (64*total - 1)/bits + 1;
   for (long i = 0; i < length; i++)
   {
      memset(poly[i], 0, sizeof(limb)*((size_t) out + 1));
      for (long k = 0; k*64 < bits && k < out; k++)
      {
         limb v = bits_at(limbs, total, (unsigned long) i*(unsigned long) bits + (unsigned long) k*64);
         long rem = bits - k*64;
         if (rem < 64) v &= ((limb) 1 << rem) - 1;
         poly[i][k] = v;
      }
   }
   return length;
}
/* FFT_combine_bits (207-267): res = sum poly[i] << (i*bits), truncated to total limbs */
void orc_combine_bits(limb *res, limb **poly, long length, long bits, long out, long total)
{
   limb *tmp = (limb *) malloc(sizeof(limb)*((size_t) out + 2));
   memset(res, 0, sizeof(limb)*(size_t) total);
   for (long i = 0; i < length; i++)
   {
      unsigned long off = (unsigned long) i*(unsigned long) bits; long q = (long)(off/64); unsigned r = (unsigned)(off % 64);
      long len = out + 1;
      if (q >= total) break;
      tmp[out + 1] = 0;
      for (long k = 0; k <= out; k++) tmp[k] = poly[i][k];
      if (r) { limb c = 0; for (long k = 0; k <= out + 1; k++) { limb v = tmp[k]; tmp[k] = (v << r) | c; c = v >> (64 - r); } len = out + 2; }
      if (q + len > total) len = total - q;
      if (add_n(res + q, res + q, tmp, len) && q + len < total) add_1(res + q + len, total - q - len, 1);
   }
   free(tmp);
}

/* --------------------------------- products ---------------------------------------------- */
static void mul_basecase(limb *r, const limb *a, const limb *b, long n)
{
   memset(r, 0, sizeof(limb)*2*(size_t) n);
   for (long i = 0; i < n; i++)
   {
      limb c = 0;
      for (long j = 0; j < n; j++) { u128 t = (u128) a[i]*b[j] + r[i + j] + c; r[i + j] = (limb) t; c = (limb)(t >> 64); }
      r[i + n] = c;
   }
}
static void mul_kara(limb *r, const limb *a, const limb *b, long n, limb *ws)
{
   if (n < 24 || (n & 1)) { mul_basecase(r, a, b, n); return; }
   long h = n/2; limb *sa = ws, *sb = ws + h, *m = ws + 2*h, *nx = ws + 4*h + 2;
   limb ca, cb; long k;
   mul_kara(r, a, b, h, nx); mul_kara(r + n, a + h, b + h, h, nx);
   ca = add_n(sa, a, a + h, h); cb = add_n(sb, b, b + h, h);
   mul_kara(m, sa, sb, h, nx); m[n] = 0;
   if (ca) m[n] += add_n(m + h, m + h, sb, h);
   if (cb) m[n] += add_n(m + h, m + h, sa, h);
   if (ca && cb) m[n] += 1;
   m[n] -= sub_n(m, m, r, n); m[n] -= sub_n(m, m, r + n, n);
   k = n + 1; if (h + k > 2*n) k = 2*n - h;
   if (add_n(r + h, r + h, m, k) && h + k < 2*n) add_1(r + h + k, 2*n - h - k, 1);
}
/* mpn_mulmod_2expp1 as used at mul_fft.c:3122 (MPIR 2.4.0; restated from its contract):
 * canonical a, b -> canonical r = a*b mod 2^(64 l)+1; r may alias a */
void orc_mulmod(limb *r, const limb *a, const limb *b, long l)
{
   if (a[l] || b[l])
   {  /* 2^NW == -1 */
      limb *z = (limb *) calloc((size_t) l + 1, sizeof(limb)), *t = (limb *) malloc(sizeof(limb)*((size_t) l + 1));
      if (a[l] && b[l]) { memset(r, 0, sizeof(limb)*((size_t) l + 1)); r[0] = 1; }
      else { mod_sub(t, z, a[l] ? b : a, l); memcpy(r, t, sizeof(limb)*((size_t) l + 1)); }
      free(z); free(t); return;
   }
   limb *p = (limb *) malloc(sizeof(limb)*2*(size_t) l), *ws = (limb *) malloc(sizeof(limb)*(8*(size_t) l + 64));
   mul_kara(p, a, b, l, ws);
   r[l] = sub_n(r, p, p + l, l) ? (limb)(slimb) -1 : 0;
   orc_normalise(r, l);
   free(p); free(ws);
}

/* new_mpn_mul (mul_fft.c:3190-3265) with the row selection of 3246 corrected to
 * revbin(s, depth + 1 - depth/2) (cf. new_mpn_mul6, 3629, 3642).  Returns 0, or -1 if the
 * parameters are illegal (the reference would segfault, 3186-3187). */
int orc_new_mpn_mul(limb *r1, const limb *i1, long n1, const limb *i2, long n2, unsigned long depth, unsigned long w)
{
   long n = 1l << depth, bits1, sq = 1l << (depth/2), j1, j2, trunc, l, size, i, j, s, t, rows2;
   limb **ii, **jj, *sa, *sb;
   if ((unsigned long) n*w % 64 || (unsigned long) n*w <= depth) return -1;
   bits1 = (long)(((unsigned long) n*w - depth)/2);
   j1 = (n1*64 - 1)/bits1 + 1; j2 = (n2*64 - 1)/bits1 + 1;
   trunc = ((j1 + j2 - 2 + 2*sq)/(2*sq))*2*sq;
   if (j1 + j2 - 1 > 2*n || trunc > 2*n) return -1;
   l = (long)((unsigned long) n*w/64); size = l + 1; rows2 = 2*n/sq;
   sa = (limb *) calloc((size_t)(2*n)*(size_t) size, sizeof(limb));
   sb = (limb *) calloc((size_t)(2*n)*(size_t) size, sizeof(limb));
   ii = (limb **) malloc(sizeof(limb *)*2*(size_t) n); jj = (limb **) malloc(sizeof(limb *)*2*(size_t) n);
   for (i = 0; i < 2*n; i++) { ii[i] = sa + i*size; jj[i] = sb + i*size; }
   orc_split_bits(ii, i1, n1, bits1, l);                       /* 3234-3236 (slab is pre-zeroed) */
   orc_mfa(0, ii, n, w, sq, trunc);                            /* 3237 */
   orc_split_bits(jj, i2, n2, bits1, l);
   orc_mfa(0, jj, n, w, sq, trunc);                            /* 3239-3242 */
   for (s = 0; s < trunc/sq; s++)                              /* 3244-3253 */
   {
      long u = (long) revbin((unsigned long) s, ilog2((unsigned long) rows2))*sq;
      for (t = 0; t < sq; t++) { j = u + t; orc_mulmod(ii[j], ii[j], jj[j], l); }
   }
   orc_mfa(1, ii, n, w, sq, trunc);                            /* 3255 */
   {
      limb *tmp = (limb *) malloc(sizeof(limb)*(size_t) size);
      for (j = 0; j < trunc; j++)                              /* 3256-3260: / 2^(depth+1) */
      { orc_mul_2exp(tmp, ii[j], l, 2*64ul*(unsigned long) l - (depth + 1)); memcpy(ii[j], tmp, sizeof(limb)*(size_t) size); }
      free(tmp);
   }
   orc_combine_bits(r1, ii, j1 + j2 - 1, bits1, l, n1 + n2);   /* 3261-3262 */
   free(ii); free(jj); free(sa); free(sb);
   return 0;
}
