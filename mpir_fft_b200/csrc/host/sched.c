/* sched.c -- data-independent schedules for the radix-2 transforms over Z/(2^NW+1).
 *
 * The reference executes its transforms recursively, one butterfly at a time, writing into two
 * scratch blocks and swapping host pointers (mul_fft.c:810-826, README:56).  On a GPU we want
 * every independent butterfly of a layer in ONE launch, so this file walks the same recursions
 * symbolically and emits ops (mfft_internal.h) instead of doing arithmetic:
 *
 *   - every logical position keeps two candidate slots (slab half 0 / half 1); an op always
 *     writes the half its target position is NOT currently in, so no op ever overwrites data
 *     another op of the same stage reads, and the pointer swaps of the reference become the
 *     index permutation `slot[]`;
 *   - the stage of an op is 1 + the latest stage that produced one of its inputs or touched
 *     one of its output slots (plain list scheduling); ops of one stage are independent.
 *
 * Everything here is plain C99 host code.  Each emitter cites the reference routine it walks.
 */
#include "sched.h"
#include <stdlib.h>
#include <string.h>

uint64_t mfft_revbin(uint64_t in, uint32_t bits)
{
   uint64_t out = 0; uint32_t i;
   for (i = 0; i < bits; i++) { out = (out << 1) | (in & 1); in >>= 1; }
   return out;
}

mfft_sched *mfft_sched_new(uint32_t S, uint64_t NW)
{
   mfft_sched *s = (mfft_sched *) calloc(1, sizeof(*s));
   uint32_t i;
   if (!s) return NULL;
   s->S = S; s->NW = NW; s->M2 = 2*NW;
   s->slot = (uint32_t *) malloc(sizeof(uint32_t) * S);
   s->wr_stage = (uint32_t *) calloc(2*(size_t)S, sizeof(uint32_t));
   s->rd_stage = (uint32_t *) calloc(2*(size_t)S, sizeof(uint32_t));
   s->cap = 1024; s->ops = (mfft_op *) malloc(sizeof(mfft_op) * s->cap);
   s->phys = (uint32_t *) malloc(sizeof(uint32_t) * S);
   s->pwr_stage = (uint32_t *) calloc((size_t) S, sizeof(uint32_t));
   s->prd_stage = (uint32_t *) calloc((size_t) S, sizeof(uint32_t));
   if (!s->slot || !s->wr_stage || !s->rd_stage || !s->ops || !s->phys || !s->pwr_stage || !s->prd_stage)
   { mfft_sched_free(s); return NULL; }
   for (i = 0; i < S; i++) { s->slot[i] = i; s->phys[i] = i; }
   return s;
}

void mfft_sched_free(mfft_sched *s)
{
   if (!s) return;
   free(s->slot); free(s->wr_stage); free(s->rd_stage); free(s->ops); free(s->stage_off);
   free(s->phys); free(s->pwr_stage); free(s->prd_stage); free(s->zero); free(s);
}

int mfft_sched_zero_from(mfft_sched *s, uint32_t first)
{
   uint32_t k;
   if (!s->zero && !(s->zero = (uint8_t *) calloc(s->S ? s->S : 1, 1))) return -1;
   for (k = first; k < s->S; k++) s->zero[k] = 1;
   return 0;
}

void mfft_sched_swap(mfft_sched *s, uint32_t a, uint32_t b)
{
   uint32_t t = s->slot[a]; s->slot[a] = s->slot[b]; s->slot[b] = t;
   t = s->phys[a]; s->phys[a] = s->phys[b]; s->phys[b] = t;
   if (s->zero) { uint8_t z = s->zero[a]; s->zero[a] = s->zero[b]; s->zero[b] = z; }
}

void mfft_sched_revbin(mfft_sched *s, uint32_t p0, uint32_t is, uint32_t bits)
{
   uint64_t j, cnt = (uint64_t)1 << bits;
   for (j = 0; j < cnt; j++)
   {
      uint64_t t = mfft_revbin(j, bits);
      if (j < t) mfft_sched_swap(s, (uint32_t)(p0 + is*j), (uint32_t)(p0 + is*t));
   }
}

/* a term of an output: sign * input * 2^(e + col*c) */
typedef struct { int sign; uint64_t e, c; } term;

static term T(int sign, uint64_t e, uint64_t c) { term t; t.sign = sign; t.e = e; t.c = c; return t; }
static const term T0 = {0, 0, 0};

static uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* outS -> position pS (from A=posA, B=posB), outT -> position pT; pT/posB may be MFFT_NONE */
static void emit(mfft_sched *s, uint32_t posA, uint32_t posB, uint32_t pS, term sa, term sb,
                 uint32_t pT, term ta, term tb)
{
   mfft_op *op; uint32_t st = 0, S = s->S;
   if (s->zero)
   {  /* known zeros: drop zero operands; an output without a term left is zero and costs nothing */
      int sz, tz;
      if (posB != MFFT_NONE && s->zero[posB]) { posB = MFFT_NONE; sb = T0; tb = T0; }
      if (s->zero[posA])
      {
         if (posB == MFFT_NONE) { s->zero[pS] = 1; if (pT != MFFT_NONE) s->zero[pT] = 1; return; }
         posA = posB; sa = sb; ta = tb; posB = MFFT_NONE; sb = T0; tb = T0;      /* only B contributes */
      }
      if (posB == MFFT_NONE) { sb = T0; tb = T0; }
      sz = (sa.sign == 0 && sb.sign == 0);
      tz = (pT == MFFT_NONE) || (ta.sign == 0 && tb.sign == 0);
      if (sz && tz) { s->zero[pS] = 1; if (pT != MFFT_NONE) s->zero[pT] = 1; return; }
      if (sz) { s->zero[pS] = 1; pS = pT; sa = ta; sb = tb; pT = MFFT_NONE; ta = T0; tb = T0; }
      else if (pT != MFFT_NONE && tz) { s->zero[pT] = 1; pT = MFFT_NONE; ta = T0; tb = T0; }
      /* "x + 0" in place: no work, but the position keeps its place in the layer structure (its
         in-place stage advances as if the op were there), so that the layers of the two halves of a
         transform stay aligned and the pass windows keep their shape */
      if (posB == MFFT_NONE && pS == posA && sa.sign == 1 && sa.e % s->M2 == 0 && sa.c % s->M2 == 0)
      {
         const uint32_t ph = s->phys[pS];
         uint32_t g = umax(s->pwr_stage[ph], s->prd_stage[ph]) + 1;
         s->pwr_stage[ph] = g; s->prd_stage[ph] = umax(s->prd_stage[ph], g);
         if (g > s->npstages) s->npstages = g;
         if (pT == MFFT_NONE) { s->zero[pS] = 0; return; }
         pS = pT; sa = ta; pT = MFFT_NONE; ta = T0;            /* only the other output is work */
      }
      s->zero[pS] = 0; if (pT != MFFT_NONE) s->zero[pT] = 0;
   }
   if (s->nops == s->cap)
   {
      s->cap *= 2; s->ops = (mfft_op *) realloc(s->ops, sizeof(mfft_op) * s->cap);
      if (!s->ops) abort();
   }
   op = &s->ops[s->nops++];
   memset(op, 0, sizeof(*op));
   op->inA = s->slot[posA];
   op->inB = (posB == MFFT_NONE) ? MFFT_NONE : s->slot[posB];
   op->outS = (s->slot[pS] < S) ? s->slot[pS] + S : s->slot[pS] - S;
   op->outT = MFFT_NONE;
   if (pT != MFFT_NONE) op->outT = (s->slot[pT] < S) ? s->slot[pT] + S : s->slot[pT] - S;
   op->sSA = (int8_t) sa.sign; op->eSA = (uint32_t)(sa.e % s->M2); op->cSA = (uint32_t)(sa.c % s->M2);
   op->sSB = (int8_t) sb.sign; op->eSB = (uint32_t)(sb.e % s->M2); op->cSB = (uint32_t)(sb.c % s->M2);
   op->sTA = (int8_t) ta.sign; op->eTA = (uint32_t)(ta.e % s->M2); op->cTA = (uint32_t)(ta.c % s->M2);
   op->sTB = (int8_t) tb.sign; op->eTB = (uint32_t)(tb.e % s->M2); op->cTB = (uint32_t)(tb.c % s->M2);
   if (op->inB == MFFT_NONE) { op->sSB = 0; op->sTB = 0; }

   st = s->wr_stage[op->inA];
   if (op->inB != MFFT_NONE) st = umax(st, s->wr_stage[op->inB]);
   st = umax(st, umax(s->rd_stage[op->outS], s->wr_stage[op->outS]));
   if (op->outT != MFFT_NONE) st = umax(st, umax(s->rd_stage[op->outT], s->wr_stage[op->outT]));
   st += 1;
   op->stage = st;
   s->rd_stage[op->inA] = umax(s->rd_stage[op->inA], st);
   if (op->inB != MFFT_NONE) s->rd_stage[op->inB] = umax(s->rd_stage[op->inB], st);
   s->wr_stage[op->outS] = st; s->slot[pS] = op->outS;
   if (op->outT != MFFT_NONE) { s->wr_stage[op->outT] = st; s->slot[pT] = op->outT; }

   /* in-place view: an op reads physical pA/pB completely, then overwrites pS/pT */
   op->pA = s->phys[posA]; op->pB = (posB == MFFT_NONE) ? MFFT_NONE : s->phys[posB];
   op->pS = s->phys[pS];   op->pT = (pT == MFFT_NONE) ? MFFT_NONE : s->phys[pT];
   st = s->pwr_stage[op->pA];
   if (op->pB != MFFT_NONE) st = umax(st, s->pwr_stage[op->pB]);
   st = umax(st, umax(s->pwr_stage[op->pS], (op->pS == op->pA || op->pS == op->pB) ? 0 : s->prd_stage[op->pS]));
   if (op->pS == op->pA || op->pS == op->pB)
   {  /* overwriting an own input: other readers of that position must be done */
      st = umax(st, s->prd_stage[op->pS]);
   }
   if (op->pT != MFFT_NONE) st = umax(st, umax(s->pwr_stage[op->pT], s->prd_stage[op->pT]));
   st += 1;
   op->pstage = st;
   s->prd_stage[op->pA] = umax(s->prd_stage[op->pA], st);
   if (op->pB != MFFT_NONE) s->prd_stage[op->pB] = umax(s->prd_stage[op->pB], st);
   s->pwr_stage[op->pS] = st;
   if (op->pT != MFFT_NONE) s->pwr_stage[op->pT] = st;
   if (st > s->npstages) s->npstages = st;
}

void mfft_sched_emit_op(mfft_sched *s, uint32_t posA, uint32_t posB,
                        uint32_t pS, int sSA, uint64_t eSA, int sSB, uint64_t eSB,
                        uint32_t pT, int sTA, uint64_t eTA, int sTB, uint64_t eTB)
{
   emit(s, posA, posB, pS, T(sSA, eSA, 0), T(sSB, eSB, 0), pT, T(sTA, eTA, 0), T(sTB, eTB, 0));
}

/* the walker's context: positions p0 + is*k, ring exponents mod M2, twist unit ws.  The column a
   batch entry stands for is cmul*col + cadd (cmul = 1, cadd = 0 except for the sqrt2 transforms,
   whose even and odd columns have different op lists and are run as two sub-batches) */
typedef struct { mfft_sched *s; uint32_t is; uint64_t ws; uint64_t cmul, cadd; } ctx;

/* a term 2^(x * column): constant part x*cadd, per-entry part x*cmul */
static uint64_t mulmod64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((unsigned __int128) a * b) % m); }
#define CT(sign, x) T((sign), mulmod64((x) % c->s->M2, c->cadd, c->s->M2), mulmod64((x) % c->s->M2, c->cmul, c->s->M2))

#define POS(k) ((uint32_t)(p0 + (uint64_t)c->is*(k)))
#define NEG(e) ((c->s->M2 - ((e) % c->s->M2)) % c->s->M2)

/* forward butterfly [a,b] -> [a+b, 2^{iw}(a-b)]   (FFT_radix2_butterfly, mul_fft.c:553-576) */
static void fwd_bfly(ctx *c, uint32_t pa, uint32_t pb, uint64_t e)
{
   emit(c->s, pa, pb, pa, T(1, 0, 0), T(1, 0, 0), pb, T(1, e, 0), T(-1, e, 0));
}

/* inverse butterfly [a,b] -> [a + 2^{-iw}b, a - 2^{-iw}b]   (mul_fft.c:639-652) */
static void inv_bfly(ctx *c, uint32_t pa, uint32_t pb, uint64_t e)
{
   emit(c->s, pa, pb, pa, T(1, 0, 0), T(1, NEG(e), 0), pb, T(1, 0, 0), T(-1, NEG(e), 0));
}

/* FFT_radix2 (786-827) / FFT_radix2_twiddle (1397-1442) */
static void fft_full(ctx *c, uint32_t p0, uint64_t n, uint64_t w, uint64_t r, uint64_t rs)
{
   uint64_t i;
   if (n == 1)
   {  /* leaf: FFT_radix2_twiddle_butterfly (517-548) with b1 = r*c*ws, b2 = (r+rs)*c*ws */
      uint64_t c1 = (r % c->s->M2) * (c->ws % c->s->M2) % c->s->M2;
      uint64_t c2 = ((r + rs) % c->s->M2) * (c->ws % c->s->M2) % c->s->M2;
      if (c1 == 0 && c2 == 0)
         emit(c->s, POS(0), POS(1), POS(0), T(1, 0, 0), T(1, 0, 0), POS(1), T(1, 0, 0), T(-1, 0, 0));
      else
      {  /* the same as an untwisted butterfly followed by two rotations: (a+b) and (a-b) are
            chunk-local in the tile executor, and each bit-granular twist then rotates one operand
            instead of two */
         emit(c->s, POS(0), POS(1), POS(0), T(1, 0, 0), T(1, 0, 0), POS(1), T(1, 0, 0), T(-1, 0, 0));
         if (c1) emit(c->s, POS(0), MFFT_NONE, POS(0), CT(1, c1), T0, MFFT_NONE, T0, T0);
         if (c2) emit(c->s, POS(1), MFFT_NONE, POS(1), CT(1, c2), T0, MFFT_NONE, T0, T0);
      }
      return;
   }
   for (i = 0; i < n; i++) fwd_bfly(c, POS(i), POS(n + i), i*w);
   fft_full(c, p0, n/2, 2*w, r, 2*rs);
   fft_full(c, POS(n), n/2, 2*w, r + rs, 2*rs);
}

/* FFT_radix2_truncate1 (1028-1074) / FFT_radix2_truncate1_twiddle (1076-1122) */
static void fft_trunc1(ctx *c, uint32_t p0, uint64_t n, uint64_t w, uint64_t r, uint64_t rs, uint64_t trunc)
{
   uint64_t i;
   if (trunc == 2*n) { fft_full(c, p0, n, w, r, rs); return; }
   if (trunc <= n)
   {
      for (i = 0; i < n; i++)      /* mpn_add_n(ii[i], ii[i], ii[i+n]), 1093-1094 */
         emit(c->s, POS(i), POS(i + n), POS(i), T(1, 0, 0), T(1, 0, 0), MFFT_NONE, T0, T0);
      fft_trunc1(c, p0, n/2, 2*w, r, 2*rs, trunc);
   } else
   {
      for (i = 0; i < n; i++) fwd_bfly(c, POS(i), POS(n + i), i*w);
      fft_full(c, p0, n/2, 2*w, r, 2*rs);
      fft_trunc1(c, POS(n), n/2, 2*w, r + rs, 2*rs, trunc - n);
   }
}

/* FFT_radix2_truncate (1128-1177) / FFT_radix2_truncate_twiddle (1179-1228): zeros past trunc */
static void fft_trunc(ctx *c, uint32_t p0, uint64_t n, uint64_t w, uint64_t r, uint64_t rs, uint64_t trunc)
{
   uint64_t i;
   if (trunc == 2*n) { fft_full(c, p0, n, w, r, rs); return; }
   if (trunc <= n) { fft_trunc(c, p0, n/2, 2*w, r, 2*rs, trunc); return; }
   for (i = 0; i < trunc - n; i++) fwd_bfly(c, POS(i), POS(n + i), i*w);
   for (i = trunc; i < 2*n; i++)   /* FFT_twiddle(ii[i], ii[i-n], i-n, n, w), 1217-1220 */
      emit(c->s, POS(i - n), MFFT_NONE, POS(i), T(1, (i - n)*w, 0), T0, MFFT_NONE, T0, T0);
   fft_full(c, p0, n/2, 2*w, r, 2*rs);
   fft_trunc1(c, POS(n), n/2, 2*w, r + rs, 2*rs, trunc - n);
}

/* IFFT_radix2 (1444-1486) / IFFT_radix2_twiddle (1964-2010) */
static void ifft_full(ctx *c, uint32_t p0, uint64_t n, uint64_t w, uint64_t r, uint64_t rs)
{
   uint64_t i;
   if (n == 1)
   {  /* FFT_radix2_twiddle_inverse_butterfly (721-752): 2^{-b1} a +- 2^{-b2} b */
      uint64_t c1 = (r % c->s->M2) * (c->ws % c->s->M2) % c->s->M2;
      uint64_t c2 = ((r + rs) % c->s->M2) * (c->ws % c->s->M2) % c->s->M2;
      /* the two rotations first, then an untwisted butterfly (see fft_full) */
      if (c1) emit(c->s, POS(0), MFFT_NONE, POS(0), CT(1, NEG(c1)), T0, MFFT_NONE, T0, T0);
      if (c2) emit(c->s, POS(1), MFFT_NONE, POS(1), CT(1, NEG(c2)), T0, MFFT_NONE, T0, T0);
      emit(c->s, POS(0), POS(1), POS(0), T(1, 0, 0), T(1, 0, 0), POS(1), T(1, 0, 0), T(-1, 0, 0));
      return;
   }
   ifft_full(c, p0, n/2, 2*w, r, 2*rs);
   ifft_full(c, POS(n), n/2, 2*w, r + rs, 2*rs);
   for (i = 0; i < n; i++) inv_bfly(c, POS(i), POS(n + i), i*w);
}

/* IFFT_radix2_truncate1 (1538-1602) / IFFT_radix2_truncate1_twiddle (1604-1668) */
static void ifft_trunc1(ctx *c, uint32_t p0, uint64_t n, uint64_t w, uint64_t r, uint64_t rs, uint64_t trunc)
{
   uint64_t i, M2 = c->s->M2;
   if (trunc == 2*n) { ifft_full(c, p0, n, w, r, rs); return; }
   if (trunc <= n)
   {
      for (i = trunc; i < n; i++)   /* (ii[i] + ii[i+n])/2, 1622-1626 */
         emit(c->s, POS(i), POS(i + n), POS(i), T(1, M2 - 1, 0), T(1, M2 - 1, 0), MFFT_NONE, T0, T0);
      ifft_trunc1(c, p0, n/2, 2*w, r, 2*rs, trunc);
      for (i = 0; i < trunc; i++)   /* mpn_addsub_n: 2*ii[i] - ii[n+i], 1630-1631 */
         emit(c->s, POS(i), POS(n + i), POS(i), T(1, 1, 0), T(-1, 0, 0), MFFT_NONE, T0, T0);
      return;
   }
   ifft_full(c, p0, n/2, 2*w, r, 2*rs);
   for (i = trunc - n; i < n; i++)  /* 1639-1647: ii[i+n] = z^i (a - b); ii[i] = 2a - b */
      emit(c->s, POS(i), POS(i + n), POS(i), T(1, 1, 0), T(-1, 0, 0),
                                     POS(i + n), T(1, i*w, 0), T(-1, i*w, 0));
   ifft_trunc1(c, POS(n), n/2, 2*w, r + rs, 2*rs, trunc - n);
   for (i = 0; i < trunc - n; i++) inv_bfly(c, POS(i), POS(n + i), i*w);
}

/* IFFT_radix2_truncate (1674-1731) / IFFT_radix2_truncate_twiddle (1733-1790) */
static void ifft_trunc(ctx *c, uint32_t p0, uint64_t n, uint64_t w, uint64_t r, uint64_t rs, uint64_t trunc)
{
   uint64_t i;
   if (trunc == 2*n) { ifft_full(c, p0, n, w, r, rs); return; }
   if (trunc <= n)
   {
      ifft_trunc(c, p0, n/2, 2*w, r, 2*rs, trunc);
      for (i = 0; i < trunc; i++)   /* doubling, 1753-1754 */
         emit(c->s, POS(i), MFFT_NONE, POS(i), T(1, 1, 0), T0, MFFT_NONE, T0, T0);
      return;
   }
   ifft_full(c, p0, n/2, 2*w, r, 2*rs);
   for (i = trunc; i < 2*n; i++)    /* FFT_twiddle(ii[i], ii[i-n], i-n, n, w), 1763-1766 */
      emit(c->s, POS(i - n), MFFT_NONE, POS(i), T(1, (i - n)*w, 0), T0, MFFT_NONE, T0, T0);
   ifft_trunc1(c, POS(n), n/2, 2*w, r + rs, 2*rs, trunc - n);
   for (i = 0; i < trunc - n; i++) inv_bfly(c, POS(i), POS(n + i), i*w);
   for (i = trunc - n; i < n; i++)  /* doubling, 1788-1789 */
      emit(c->s, POS(i), MFFT_NONE, POS(i), T(1, 1, 0), T0, MFFT_NONE, T0, T0);
}

/* FFT_radix2_negacyclic, even w (1346-1373): twist by 2^{i*w/2}, then one butterfly layer */
static void fft_negacyclic(ctx *c, uint32_t p0, uint64_t n, uint64_t w)
{
   uint64_t i, h = w/2;
   for (i = 0; i < n; i++)
   {
      uint64_t ea = i*h, eb = (n + i)*h, e = i*w;
      emit(c->s, POS(i), POS(n + i), POS(i), T(1, ea, 0), T(1, eb, 0),
                                     POS(n + i), T(1, ea + e, 0), T(-1, eb + e, 0));
   }
   fft_full(c, p0, n/2, 2*w, 0, 0);
   fft_full(c, POS(n), n/2, 2*w, 0, 0);
}

/* IFFT_radix2_negacyclic, even w (1882-1886, 1940-1960) */
static void ifft_negacyclic(ctx *c, uint32_t p0, uint64_t n, uint64_t w)
{
   uint64_t i, h = w/2;
   ifft_full(c, p0, n/2, 2*w, 0, 0);
   ifft_full(c, POS(n), n/2, 2*w, 0, 0);
   for (i = 0; i < n; i++)
   {  /* s = a + 2^{-iw} b, t = a - 2^{-iw} b, then s *= 2^{-i w/2}, t *= 2^{-(n+i) w/2} */
      uint64_t ua = NEG(i*h), ub = NEG((n + i)*h), e = NEG(i*w);
      emit(c->s, POS(i), POS(n + i), POS(i), T(1, ua, 0), T(1, ua + e, 0),
                                     POS(n + i), T(1, ub, 0), T(-1, ub + e, 0));
   }
}


/* ---- the sqrt2 transforms: length 4n over Z/(2^(nw)+1) with the 4n-th root of unity
 *      z1 = sqrt2^w,  sqrt2 = 2^(3nw/4) - 2^(nw/4)   (mul_fft.c:578-617, 959-1021) --------------------
 * z1^j for even j is the power of two 2^((j/2) w); for odd j (w odd) it is
 *      z1^j = 2^e (2^(nw/2) - 1),   e = (j w - 1)/2 + nw/4,
 * i.e. one rotation followed by "x 2^(nw/2) - x", which is chunk-aligned.  The column schedule below
 * covers positions 0..2 n2-1 of one MFA column (first half rows, then second half rows): the
 * outermost layer between the halves (2240-2262 / 2700-2731), then the two column transforms.
 * j = column + row*n1 has the parity of the column (n1 is even), so the op list depends on the
 * column's parity: par = 0 / 1 builds the list of the even / odd columns (column = 2 col + par),
 * par = -1 (even w: the root is the power of two 2^(w/2), 2264-2279) one list for all columns. */
static void sq2_apply(ctx *c, uint32_t p)          /* x <- x 2^(NW/2) - x, in place */
{
   emit(c->s, p, p, p, T(1, c->s->NW/2, 0), T(-1, 0, 0), MFFT_NONE, T0, T0);
}

/* position dst <- src * z1^(+-j), j = column + row*n1, as one rotation (+ sq2_apply for odd j) */
static void sqrt2_twiddle(ctx *c, uint32_t src, uint32_t dst, uint64_t row, uint64_t n1, uint64_t w, int par, int inverse, int pad)
{
   const uint64_t NW = c->s->NW, M2 = c->s->M2;
   if (par < 0)
   {  /* 2^(j w/2): constant part row*n1*w/2, per-column part w/2 */
      uint64_t e0 = mulmod64(row % M2, (n1*(w/2)) % M2, M2), ec = (w/2) % M2;
      if (inverse) { e0 = (M2 - e0) % M2; ec = (M2 - ec) % M2; }
      emit(c->s, src, MFFT_NONE, dst, T(1, e0, ec), T0, MFFT_NONE, T0, T0);
      return;
   }
   if (par == 0)
   {  /* j even: 2^((j/2) w) = 2^(col w + row (n1/2) w) */
      uint64_t e0 = mulmod64(row % M2, ((n1/2)*w) % M2, M2), ec = w % M2;
      if (inverse) { e0 = (M2 - e0) % M2; ec = (M2 - ec) % M2; }
      emit(c->s, src, MFFT_NONE, dst, T(1, e0, ec), T0, MFFT_NONE, T0, T0);
      /* ping-pong view only: a copy where the odd columns apply (2^(NW/2) - 1), so that both classes
         leave every position in the same slab half */
      if (pad) emit(c->s, dst, MFFT_NONE, dst, T(1, 0, 0), T0, MFFT_NONE, T0, T0);
      return;
   }
   {  /* j odd: (j w - 1)/2 = col w + (w-1)/2 + row (n1/2) w.  Forward exponent e = that + NW/4; the
         inverse is z1^-j = 2^(2NW - (j w - 1)/2 - 1 + NW/4) (2^(NW/2) - 1)   (mul_fft.c:653-672) */
      uint64_t h = (mulmod64(row % M2, ((n1/2)*w) % M2, M2) + (w - 1)/2) % M2, ec = w % M2, e0;
      if (!inverse) e0 = (h + NW/4) % M2;
      else { e0 = (2*M2 - h - 1 + NW/4) % M2; ec = (M2 - ec) % M2; }
      emit(c->s, src, MFFT_NONE, dst, T(1, e0, ec), T0, MFFT_NONE, T0, T0);
      sq2_apply(c, dst);
   }
}

/* position p <- p * z1^j, z1 = sqrt2^w the 4n-th root of unity (w odd), 0 <= j < 4n:
   even j a power of two (FFT_twiddle, 926), odd j FFT_twiddle_sqrt2 (972) */
static void z1pow(ctx *c, uint32_t p, uint64_t j, uint64_t w)
{
   const uint64_t NW = c->s->NW, M2 = c->s->M2;
   if (!(j & 1)) { const uint64_t e = mulmod64((j/2) % M2, w % M2, M2); if (e) emit(c->s, p, MFFT_NONE, p, T(1, e, 0), T0, MFFT_NONE, T0, T0); return; }
   emit(c->s, p, MFFT_NONE, p, T(1, ((mulmod64(j % (2*M2), w % (2*M2), 2*M2) - 1)/2 + NW/4) % M2, 0), T0, MFFT_NONE, T0, T0);
   sq2_apply(c, p);
}

/* FFT_radix2_negacyclic / IFFT_radix2_negacyclic for odd w (1301-1343, 1892-1938): the twist by z1^p
   needs sqrt2 at the odd positions */
static void fft_negacyclic_odd(ctx *c, uint32_t p0, uint64_t n, uint64_t w)
{
   uint64_t i;
   for (i = 0; i < n; i++)
   {
      z1pow(c, POS(i), i, w); z1pow(c, POS(n + i), n + i, w);
      fwd_bfly(c, POS(i), POS(n + i), i*w);
   }
   fft_full(c, p0, n/2, 2*w, 0, 0);
   fft_full(c, POS(n), n/2, 2*w, 0, 0);
}

static void ifft_negacyclic_odd(ctx *c, uint32_t p0, uint64_t n, uint64_t w)
{
   uint64_t i;
   ifft_full(c, p0, n/2, 2*w, 0, 0);
   ifft_full(c, POS(n), n/2, 2*w, 0, 0);
   for (i = 0; i < n; i++)
   {
      inv_bfly(c, POS(i), POS(n + i), i*w);
      z1pow(c, POS(i), (4*n - i) % (4*n), w); z1pow(c, POS(n + i), 3*n - i, w);
   }
}

/* FFT_radix2_sqrt2 / FFT_radix2_truncate_sqrt2 / IFFT_radix2_sqrt2 / IFFT_radix2_truncate_sqrt2
   (mul_fft.c:839, 1230, 1488, 1792): length 4n over positions 0..4n-1 (stride 1), root z1 = sqrt2^w.
   trunc in (2n, 4n], even; outputs in the bit-reversed order of the two halves, like FFT_radix2. */
int mfft_sched_emit_sqrt2_1d(mfft_sched *s, int inverse, uint64_t n, uint64_t w, uint64_t trunc)
{
   ctx cc, *c = &cc; uint64_t j, NW = s->NW, tr2; uint32_t p0 = 0;
   c->s = s; c->is = 1; c->ws = 0; c->cmul = 1; c->cadd = 0;
   if (n < 2 || (n & (n - 1)) || n*w != NW || s->S != 4*n || trunc <= 2*n || trunc > 4*n || (trunc & 1)) return -1;
   if (!(w & 1))
   {  /* the root is the power of two 2^(w/2): the plain transforms of length 2*(2n) (851-856, 1244-1248) */
      if (trunc == 4*n) return mfft_sched_emit(s, inverse ? MFFT_T_IFFT : MFFT_T_FFT, 0, 1, 2*n, w/2, 0, 0, 0, 0);
      return mfft_sched_emit(s, inverse ? MFFT_T_IFFT_TRUNC : MFFT_T_FFT_TRUNC, 0, 1, 2*n, w/2, 0, 0, 0, trunc);
   }
   if (NW % 4) return -1;
   tr2 = trunc - 2*n;
   if (!inverse)
   {
      for (j = 0; j < tr2; j++)
      {
         emit(s, (uint32_t) j, (uint32_t)(2*n + j), (uint32_t) j, T(1, 0, 0), T(1, 0, 0), (uint32_t)(2*n + j), T(1, 0, 0), T(-1, 0, 0));
         z1pow(c, (uint32_t)(2*n + j), j, w);
      }
      for (; j < 2*n; j++)
      {  /* the partner is zero: second half = z1^j * first half (1271-1278) */
         emit(s, (uint32_t) j, MFFT_NONE, (uint32_t)(2*n + j), T(1, 0, 0), T0, MFFT_NONE, T0, T0);
         z1pow(c, (uint32_t)(2*n + j), j, w);
      }
      fft_full(c, p0, n, w, 0, 0);
      { uint32_t p0 = (uint32_t)(2*n); if (tr2 == 2*n) fft_full(c, p0, n, w, 0, 0); else fft_trunc1(c, p0, n, w, 0, 0, tr2); }
   } else
   {
      ifft_full(c, p0, n, w, 0, 0);
      for (j = tr2; j < 2*n; j++)
      {
         emit(s, (uint32_t) j, MFFT_NONE, (uint32_t)(2*n + j), T(1, 0, 0), T0, MFFT_NONE, T0, T0);
         z1pow(c, (uint32_t)(2*n + j), j, w);
      }
      { uint32_t p0 = (uint32_t)(2*n); if (tr2 == 2*n) ifft_full(c, p0, n, w, 0, 0); else ifft_trunc1(c, p0, n, w, 0, 0, tr2); }
      for (j = 0; j < tr2; j++)
      {
         z1pow(c, (uint32_t)(2*n + j), (4*n - j) % (4*n), w);
         emit(s, (uint32_t) j, (uint32_t)(2*n + j), (uint32_t) j, T(1, 0, 0), T(1, 0, 0), (uint32_t)(2*n + j), T(1, 0, 0), T(-1, 0, 0));
      }
      for (; j < 2*n; j++) emit(s, (uint32_t) j, MFFT_NONE, (uint32_t) j, T(1, 1, 0), T0, MFFT_NONE, T0, T0);
   }
   return 0;
}

int mfft_sched_emit_sqrt2_cols(mfft_sched *s, int inverse, uint64_t n2, uint64_t n1, uint64_t w, uint64_t trunc2, int par, int pad)
{
   ctx cc, *c = &cc; uint64_t r, NW = s->NW, n = NW/w; uint32_t depth = 0, p0 = 0;
   (void) p0;
   c->s = s; c->is = 1; c->ws = w;
   c->cmul = (par < 0) ? 1 : 2; c->cadd = (par < 0) ? 0 : (uint64_t) par;
   if (n*w != NW || 2*n != n1*n2 || s->S != 2*n2 || (n1 & 1) || NW % 4) return -1;
   if ((par < 0) != ((w & 1) == 0)) return -1;              /* parity classes exist iff w is odd */
   if (trunc2 < 2 || trunc2 > n2 || (trunc2 & 1)) return -1;
   while (((uint64_t)1 << depth) < n2) depth++;
   if (!inverse)
   {
      for (r = 0; r < trunc2; r++)
      {  /* [a, b] -> [a + b, z1^j (a - b)]   (FFT_radix2_butterfly(_sqrt2), 2244-2262) */
         emit(s, (uint32_t) r, (uint32_t)(n2 + r), (uint32_t) r, T(1, 0, 0), T(1, 0, 0), (uint32_t)(n2 + r), T(1, 0, 0), T(-1, 0, 0));
         sqrt2_twiddle(c, (uint32_t)(n2 + r), (uint32_t)(n2 + r), r, n1, w, par, 0, pad);
      }
      for (; r < n2; r++)     /* b = 0: second half row = z1^j a   (FFT_twiddle(_sqrt2), 2265-2271) */
         sqrt2_twiddle(c, (uint32_t) r, (uint32_t)(n2 + r), r, n1, w, par, 0, pad);
      /* the two column transforms with the z^(r c) twist (2283, 2337), then the row relabels */
      { uint32_t p0 = 0; fft_full(c, p0, n2/2, w*n1, 0, 1); }
      mfft_sched_revbin(s, 0, 1, depth);
      { uint32_t p0 = (uint32_t) n2; fft_trunc1(c, p0, n2/2, w*n1, 0, 1, trunc2); }
      mfft_sched_revbin(s, (uint32_t) n2, 1, depth);
   } else
   {
      uint64_t j;
      mfft_sched_revbin(s, 0, 1, depth);
      { uint32_t p0 = 0; ifft_full(c, p0, n2/2, w*n1, 0, 1); }                                       /* 2629-2645 */
      for (j = 0; j < trunc2; j++)                                                                   /* 2672-2681 */
      {
         uint64_t t = mfft_revbin(j, depth);
         if (j < t) mfft_sched_swap(s, (uint32_t)(n2 + j), (uint32_t)(n2 + t));
      }
      for (r = trunc2; r < n2; r++)                                                                  /* 2683-2694 */
         sqrt2_twiddle(c, (uint32_t) r, (uint32_t)(n2 + r), r, n1, w, par, 0, pad);
      { uint32_t p0 = (uint32_t) n2; ifft_trunc1(c, p0, n2/2, w*n1, 0, 1, trunc2); }                  /* 2698 */
      for (r = 0; r < trunc2; r++)
      {  /* [a, b] -> [a + z1^-j b, a - z1^-j b]   (2702-2745) */
         sqrt2_twiddle(c, (uint32_t)(n2 + r), (uint32_t)(n2 + r), r, n1, w, par, 1, pad);
         emit(s, (uint32_t) r, (uint32_t)(n2 + r), (uint32_t) r, T(1, 0, 0), T(1, 0, 0), (uint32_t)(n2 + r), T(1, 0, 0), T(-1, 0, 0));
      }
      for (; r < n2; r++)     /* doubling of the first-half rows without a partner (2747-2748) */
         emit(s, (uint32_t) r, MFFT_NONE, (uint32_t) r, T(1, 1, 0), T0, MFFT_NONE, T0, T0);
   }
   return 0;
}

int mfft_sched_emit(mfft_sched *s, mfft_transform_kind kind, uint32_t p0, uint32_t is,
                    uint64_t n, uint64_t w, uint64_t ws, uint64_t r, uint64_t rs, uint64_t trunc)
{
   ctx c; c.s = s; c.is = is; c.ws = ws; c.cmul = 1; c.cadd = 0;
   if (n == 0 || (n & (n - 1)) || n*w != s->NW) return -1;
   if ((uint64_t) p0 + (uint64_t) is*(2*n - 1) >= s->S) return -1;
   switch (kind)
   {
   case MFFT_T_FFT:  fft_full(&c, p0, n, w, r, rs); break;
   case MFFT_T_IFFT: ifft_full(&c, p0, n, w, r, rs); break;
   case MFFT_T_FFT_TRUNC: case MFFT_T_FFT_TRUNC1: case MFFT_T_IFFT_TRUNC: case MFFT_T_IFFT_TRUNC1:
      /* odd trunc recurses forever in the reference (1438); reject it */
      if (trunc < 2 || trunc > 2*n || (trunc & 1)) return -1;
      if (kind == MFFT_T_FFT_TRUNC)   fft_trunc(&c, p0, n, w, r, rs, trunc);
      if (kind == MFFT_T_FFT_TRUNC1)  fft_trunc1(&c, p0, n, w, r, rs, trunc);
      if (kind == MFFT_T_IFFT_TRUNC)  ifft_trunc(&c, p0, n, w, r, rs, trunc);
      if (kind == MFFT_T_IFFT_TRUNC1) ifft_trunc1(&c, p0, n, w, r, rs, trunc);
      break;
   case MFFT_T_FFT_NEGACYCLIC:
      if (n < 2 || ((w & 1) && (s->NW % 4))) return -1;
      if (w & 1) fft_negacyclic_odd(&c, p0, n, w); else fft_negacyclic(&c, p0, n, w);
      break;
   case MFFT_T_IFFT_NEGACYCLIC:
      if (n < 2 || ((w & 1) && (s->NW % 4))) return -1;
      if (w & 1) ifft_negacyclic_odd(&c, p0, n, w); else ifft_negacyclic(&c, p0, n, w);
      break;
   default: return -1;
   }
   return 0;
}

int mfft_sched_finish(mfft_sched *s)
{
   size_t i; uint32_t st, maxst = 0; mfft_op *sorted; uint32_t *cnt;
   for (i = 0; i < s->nops; i++) if (s->ops[i].stage > maxst) maxst = s->ops[i].stage;
   s->nstages = maxst;
   free(s->stage_off);
   s->stage_off = (uint32_t *) calloc((size_t) maxst + 2, sizeof(uint32_t));
   cnt = (uint32_t *) calloc((size_t) maxst + 2, sizeof(uint32_t));
   sorted = (mfft_op *) malloc(sizeof(mfft_op) * (s->nops ? s->nops : 1));
   if (!s->stage_off || !cnt || !sorted) { free(cnt); free(sorted); return -1; }
   for (i = 0; i < s->nops; i++) cnt[s->ops[i].stage]++;
   /* stage k (1-based) occupies [stage_off[k-1], stage_off[k]) */
   s->stage_off[0] = 0;
   for (st = 1; st <= maxst; st++) s->stage_off[st] = s->stage_off[st - 1] + cnt[st];
   memset(cnt, 0, sizeof(uint32_t) * ((size_t) maxst + 2));
   for (i = 0; i < s->nops; i++)
   {
      st = s->ops[i].stage;
      sorted[s->stage_off[st - 1] + cnt[st]++] = s->ops[i];
   }
   free(s->ops); s->ops = sorted; s->cap = s->nops ? s->nops : 1;
   free(cnt);
   return 0;
}
