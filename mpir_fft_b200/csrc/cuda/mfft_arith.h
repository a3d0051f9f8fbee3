/* mfft_arith.h -- lane-local arithmetic mod p = 2^NW + 1 (NW = 64*l), host+device.
 *
 * Everything a transform needs is one primitive:
 *
 *        out  =  sA * A * 2^eA  +  sB * B * 2^eB      (mod p),   sA, sB in {+1, -1, 0}
 *
 * (butterflies mul_fft.c:517-576, 639-752; FFT_twiddle 926-957; the add / 2a-b / halving steps
 * of the truncated transforms 1093, 1624-1631, 1641-1643).  A coefficient is `l` limbs of body
 * plus a signed top limb: value = body + top*2^NW == body - top.
 *
 * Multiplying the body by 2^e, e = 64*y + bs < NW, is a negacyclic bit rotation: bits
 * [0, NW-e) move up by e, bits [NW-e, NW) wrap to [0, e) with their sign flipped.  A negated
 * bit field F over [a,b) is written as its complement plus two constants,
 *        -F = ~F|[a,b) + 2^a - 2^b,
 * so every output limb is a plain sum of two (possibly complemented) rotated input limbs:
 *
 *   term +X*2^e:  complement bits [0,e);   constants  +1 at bit 0,   -(1+topX)*2^e at bit e
 *   term -X*2^e:  complement bits [e,NW);  constants  +(1+topX)*2^e at bit e,   -2^NW
 *   (e == 0:  +X keeps topX in the output top;  -X: +1 at bit 0 and top -= 1 + topX)
 *
 * "+1 at bit 0" is the carry-in of the add chain (a second one is folded into the top limb
 * because 1 == -2^NW); "-2^NW" is top -= 1; the constant at bit e is injected into the
 * stored result by one lane (mfft_inject) -- it is (topB - topA)*2^bs for a butterfly, so its
 * ripple is almost always one limb.
 *
 * The add chain itself is carry-lookahead across the 32 lanes of a warp: each lane adds its M
 * limbs with carry-in 0 and reports generate/propagate; mfft_lookahead turns the two ballots
 * into every lane's carry-in with one 64-bit addition.
 */
#ifndef MFFT_ARITH_H
#define MFFT_ARITH_H

#include <stdint.h>

#ifdef __CUDACC__
#define MFFT_HD __host__ __device__ __forceinline__
#else
#define MFFT_HD static inline
#endif

typedef uint64_t mfft_limb;
typedef __int128 mfft_i128;
typedef unsigned __int128 mfft_u128;

/* one term of a lincomb, pre-digested for the lane loop */
typedef struct {
   const mfft_limb *p;  /* block (body limbs 0..l-1, top at [l]) */
   uint32_t y;          /* limb rotation */
   uint32_t bs;         /* bit shift 0..63 */
   int      neg;        /* 1: the whole term is subtracted (after folding e >= NW) */
   int      present;
   uint32_t ps;         /* storage index of limb q is q + (q >> ps): 31 = dense (global memory),
                           log2(M) = one pad limb per M limbs (bank-conflict-free shared memory) */
   mfft_i128 K;         /* constant to inject at limb y (0 if none) */
} mfft_term;

/* fold exponent e (mod 2*NW) and sign into (y, bs, neg); returns the term's contributions to
 * the output top limb in *top_acc and the number of "+1 at bit 0" constants in *ones. */
MFFT_HD void mfft_term_setup2(mfft_term *t, const mfft_limb *blk, int64_t top, uint32_t ps, uint32_t l,
                              int sign, uint64_t e, int64_t *top_acc, int *ones)
{
   uint64_t NW = 64ull * l;
   t->p = blk; t->present = (sign != 0); t->K = 0; t->y = 0; t->bs = 0; t->neg = 0; t->ps = ps;
   if (!sign) return;
   if (e >= NW) { e -= NW; sign = -sign; }
   t->y = (uint32_t)(e >> 6); t->bs = (uint32_t)(e & 63); t->neg = (sign < 0);
   {
      if (e == 0)
      {
         if (sign > 0) *top_acc += top;
         else { *ones += 1; *top_acc -= 1 + top; }
      } else
      {
         mfft_i128 k = ((mfft_i128)(1 + top)) * ((mfft_i128)1 << t->bs);
         if (sign > 0) { *ones += 1; t->K = -k; }
         else          { *top_acc -= 1; t->K = k; }
      }
   }
}

/* dense blocks with the top limb stored at blk[l] (global memory layout) */
MFFT_HD void mfft_term_setup(mfft_term *t, const mfft_limb *blk, uint32_t l, int sign, uint64_t e,
                             int64_t *top_acc, int *ones)
{
   mfft_term_setup2(t, blk, sign ? (int64_t) blk[l] : 0, 31, l, sign, e, top_acc, ones);
}

/* limb k of the complemented rotation of the term's body */
MFFT_HD mfft_limb mfft_term_limb(const mfft_term *t, uint32_t l, uint32_t k)
{
   uint32_t q = (k >= t->y) ? k - t->y : k + l - t->y;
   mfft_limb v = t->p[q + (q >> t->ps)], m;
   if (t->bs)
   {
      uint32_t q1 = (q == 0) ? l - 1 : q - 1;
      v = (v << t->bs) | (t->p[q1 + (q1 >> t->ps)] >> (64 - t->bs));
   }
   m = (k < t->y) ? ~(mfft_limb)0 : ((k == t->y) ? (((mfft_limb)1 << t->bs) - 1) : 0);
   if (t->neg) m = ~m;
   return v ^ m;
}

/* carry-lookahead: G, P = ballots of generate / propagate per lane, cin = carry into lane 0.
 * Returns the carry-in of every lane in bits 0..31 and the carry out of lane 31 in bit 32. */
MFFT_HD uint64_t mfft_lookahead(uint32_t G, uint32_t P, uint32_t cin)
{
   uint64_t a = G, b = (uint64_t)(G | P);
   uint64_t s = a + b + cin;
   return s ^ a ^ b;
}

/* add the signed 128-bit constant K at limb y of the stored block `out`, rippling upwards;
 * whatever leaves limb l-1 goes into *top. */
MFFT_HD void mfft_inject2(mfft_limb *out, uint32_t ps, uint32_t l, int64_t *top, uint32_t y, mfft_i128 K)
{
   uint32_t pos = y;
   mfft_limb v, nv; int64_t carry;
   if (K == 0) return;
   v = out[pos + (pos >> ps)]; nv = v + (mfft_limb) K;
   carry = (int64_t)(K >> 64) + (nv < v ? 1 : 0);
   out[pos + (pos >> ps)] = nv; pos++;
   while (carry != 0)
   {
      if (pos == l) { *top += carry; return; }
      v = out[pos + (pos >> ps)]; nv = v + (mfft_limb) carry;
      carry = (carry > 0) ? (nv < v ? 1 : 0) : (nv > v ? -1 : 0);
      out[pos + (pos >> ps)] = nv; pos++;
   }
}
MFFT_HD void mfft_inject(mfft_limb *out, uint32_t l, int64_t *top, uint32_t y, mfft_i128 K)
{
   mfft_inject2(out, 31, l, top, y, K);
}

#endif
