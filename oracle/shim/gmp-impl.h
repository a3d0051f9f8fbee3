/* oracle/shim/gmp-impl.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for MPIR's internal header: the macros and the MPIR-only mpn
 * functions that /root/reference/mul_fft.c calls, restated in portable C on
 * top of the GMP 6.3.0 runtime.  Call sites in the reference:
 *   mpn_sumdiff_n        mul_fft.c:312-460 (26 sites)
 *   mpn_addsub_n         mul_fft.c:1565, 1631
 *   mpn_neg_n            mul_fft.c:332 ... (30 sites; MPIR tolerates n==0)
 *   mpn_mulmod_2expp1    mul_fft.c:3062, 3122, 3138, 4256
 *   mpn_mul_fft_aux, mpn_fft_best_k, mpn_fft_next_size: only reached from the
 *   out-of-scope new_mpn_mul3/4/5/6 and the _combined IFFT; stubbed.
 */
#ifndef ORACLE_SHIM_GMP_IMPL_H
#define ORACLE_SHIM_GMP_IMPL_H
#include "mpir.h"

#define CNST_LIMB(c) ((mp_limb_t)(c##UL))
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#define MPN_COPY(d, s, n) memmove((d), (s), (size_t)(n) * sizeof(mp_limb_t))
#define MPN_ZERO(d, n)    memset((d), 0, (size_t)(n) * sizeof(mp_limb_t))

/* TMP_* on malloc with a per-function free list */
struct shim_tmp_node { struct shim_tmp_node *next; };
static inline void *shim_tmp_alloc(struct shim_tmp_node **head, size_t bytes)
{
   struct shim_tmp_node *nd = (struct shim_tmp_node *) malloc(sizeof(*nd) + 16 + bytes);
   if (!nd) abort();
   nd->next = *head; *head = nd;
   return (char *) nd + 16;
}
static inline void shim_tmp_free(struct shim_tmp_node **head)
{
   while (*head) { struct shim_tmp_node *nx = (*head)->next; free(*head); *head = nx; }
}
#define TMP_DECL  struct shim_tmp_node *__shim_tmp = NULL
#define TMP_MARK  do { } while (0)
#define TMP_BALLOC_LIMBS(n) ((mp_limb_t *) shim_tmp_alloc(&__shim_tmp, (size_t)(n) * sizeof(mp_limb_t)))
#define TMP_FREE  shim_tmp_free(&__shim_tmp)

/* s = a + b, d = a - b over n limbs; returns 2*carry + borrow.
   Outputs may alias inputs (limb-wise read-before-write). */
static inline mp_limb_t mpn_sumdiff_n(mp_ptr s, mp_ptr d, mp_srcptr a, mp_srcptr b, mp_size_t n)
{
   mp_limb_t cy = 0, bw = 0; mp_size_t i;
   for (i = 0; i < n; i++)
   {
      mp_limb_t x = a[i], y = b[i];
      unsigned __int128 t = (unsigned __int128) x + y + cy;
      mp_limb_t df = x - y - bw;
      bw = (x < y) || (x == y && bw);
      s[i] = (mp_limb_t) t; cy = (mp_limb_t)(t >> 64);
      d[i] = df;
   }
   return 2*cy + bw;
}

/* r = a + b - c over n limbs; returns carry - borrow (signed, as a limb) */
static inline mp_limb_t mpn_addsub_n(mp_ptr r, mp_srcptr a, mp_srcptr b, mp_srcptr c, mp_size_t n)
{
   __int128 acc = 0; mp_size_t i;
   for (i = 0; i < n; i++)
   {
      acc += (__int128)(unsigned __int128) a[i];
      acc += (__int128)(unsigned __int128) b[i];
      acc -= (__int128)(unsigned __int128) c[i];
      r[i] = (mp_limb_t) acc;
      acc >>= 64;   /* arithmetic */
   }
   return (mp_limb_t) acc;
}

static inline mp_limb_t mpn_neg_n(mp_ptr r, mp_srcptr s, mp_size_t n)
{
   if (n <= 0) return 0;
   return __gmpn_neg(r, s, n);
}

static inline void mpn_copyi(mp_ptr d, mp_srcptr s, mp_size_t n) { MPN_COPY(d, s, n); }

static inline void mpn_rrandom(mp_ptr r, gmp_randstate_t st, mp_size_t n)
{
   (void) st; if (n > 0) __gmpn_random2(r, n);
}

static inline void mpn_urandomb(mp_ptr r, gmp_randstate_t st, mp_bitcnt_t bits)
{
   mpz_t z; mp_size_t n = (mp_size_t)((bits + 63)/64), k, i;
   __gmpz_init(z); __gmpz_urandomb(z, st, bits);
   k = z->_mp_size; if (k < 0) k = -k;
   for (i = 0; i < n; i++) r[i] = (i < k) ? z->_mp_d[i] : 0;
   __gmpz_clear(z);
}

/* r = i1*i2 mod 2^bits+1 (bits a multiple of 64 on every reference path).
   c bit0 <=> i1 == 2^bits, c bit1 <=> i2 == 2^bits (mul_fft.c:3250, 3061).
   Returns 1 iff the result is 2^bits (limbs of r then zero). tt: 2*l limbs. */
static inline mp_limb_t mpn_mulmod_2expp1(mp_ptr r, mp_srcptr i1, mp_srcptr i2,
                                          mp_limb_t c, mp_limb_t bits, mp_ptr tt)
{
   mp_size_t l = (mp_size_t)(bits/64), k;
   if (c & 1)
   {
      if (c & 2) { MPN_ZERO(r, l); r[0] = 1; return 0; }   /* (-1)(-1) = 1 */
      i1 = i2;                                              /* -i2 */
   }
   if (c & 3)
   {  /* r = -i1 mod p = p - i1 for i1 != 0 */
      int zero = 1;
      for (k = 0; k < l; k++) if (i1[k]) { zero = 0; break; }
      if (zero) { MPN_ZERO(r, l); return 0; }
      __gmpn_neg(r, i1, l);                  /* 2^bits - i1 */
      return __gmpn_add_1(r, r, l, 1);       /* +1; carry <=> result 2^bits */
   }
   __gmpn_mul_n(tt, i1, i2, l);
   {
      mp_limb_t bw = __gmpn_sub_n(r, tt, tt + l, l);
      if (bw) return __gmpn_add_1(r, r, l, 1);
   }
   return 0;
}

/* out-of-scope MPIR FFT entry points (only new_mpn_mul3/4/5/6 reach them) */
static inline mp_limb_t mpn_mul_fft_aux(mp_ptr r, mp_size_t l, mp_srcptr a, mp_size_t an,
                                        mp_srcptr b, mp_size_t bn, int k, int flag)
{
   /* MPIR contract (mul_fft.c:3398): {r,l} = a*b mod 2^(64l)+1, returns the top bit;
      r may alias a.  Restated as a plain product followed by the lo - hi fold. */
   mp_limb_t top = 0;
   mp_limb_t *t = (mp_limb_t *) calloc((size_t)(2*l + 2), sizeof(mp_limb_t));
   (void) k; (void) flag;
   if (an >= bn) __gmpn_mul(t, a, an, b, bn); else __gmpn_mul(t, b, bn, a, an);
   if (__gmpn_sub_n(r, t, t + l, l)) top = __gmpn_add_1(r, r, l, 1);
   free(t);
   return top;
}
static inline int mpn_fft_best_k(mp_size_t n, int sqr) { (void) n; (void) sqr; return 4; }
static inline mp_size_t mpn_fft_next_size(mp_size_t pl, int k)
{
   mp_size_t m = (mp_size_t)1 << k; return ((pl + m - 1)/m)*m;
}

#endif
