/* api.c -- the remaining reference entry points (host pointer tables in, host blocks out).
 *
 * The reference's transforms take `mp_limb_t **ii`: a table of host pointers into a slab of
 * (l+1)-limb blocks which they permute (README:56).  Here each call gathers the blocks into one
 * pinned-size staging buffer, copies it to an HBM slab, runs the device schedule, and writes result
 * k back into the block ii[k] points to (the identity permutation, which the reference's contract
 * allows: after return {ii[*], *t1, *t2} is still a permutation of the caller's block set).
 * t1/t2/temp are accepted and never touched.  These per-call entry points exist for drop-in and
 * parity testing; the fast path is the device-resident plan API (mul.c).
 */
#define _POSIX_C_SOURCE 200809L
#include "runtime.h"
#include "../../../include/mpirfft_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

mp_limb_t mpir_revbin(mp_limb_t in, mp_bitcnt_t bits) { return (mp_limb_t) mfft_revbin(in, (uint32_t) bits); }

/* Run a schedule over S host blocks.  in[k] == NULL: position k starts as zero.  out[k] == NULL:
 * position k is not written back.  Takes ownership of s. */
/* device (and host staging) scratch of the per-call entry points: one grow-only arena, so that a caller
   which drives the single-block primitives in a loop (the reference's own tests do, thousands of
   times) does not pay an allocation per call.  Used under the library lock only. */
static unsigned char *g_arena = NULL; static size_t g_arena_bytes = 0;
static limb_t *g_stage = NULL; static size_t g_stage_bytes = 0;

static void *arena_take(size_t *off, size_t bytes)
{
   void *p = g_arena + *off;
   *off += (bytes + 255) & ~(size_t) 255;
   return p;
}

static void run_on_host_blocks(const char *fn, mfft_sched *s, uint32_t l, mp_limb_t **in, mp_limb_t **out,
                               uint32_t col, int normalise)
{
   /* One packed upload per call: [batch | dst base | ops | moves | slab half 0] lie in this order both in
      the pinned host staging buffer and in the device arena, followed on the device by slab half 1 and the
      result blocks.  A call is one copy in, the schedule's launches, one copy out. */
   uint32_t S = s->S, k, pitch = l + 1;
   size_t half = (size_t) S * pitch * sizeof(limb_t), need, off = 0, nops, o_b, o_base, o_ops, o_mv, o_slab, up;
   unsigned char *st; limb_t *stage, *d_slab, *d_dst; mfft_batch *b, *d_b; mfft_move *mv, *d_mv;
   uint32_t *d_base; mfft_dsched ds; mfft_geom g;
   mfft_lock();
   mfft_require_device(fn);
   if (mfft_sched_finish(s) != 0) mfft_die(fn, "out of host memory");
   nops = s->nops ? s->nops : 1;
   o_b = off;    (void) arena_take(&off, sizeof *b);
   o_base = off; (void) arena_take(&off, sizeof(uint32_t));
   o_ops = off;  (void) arena_take(&off, sizeof(mfft_op)*nops);
   o_mv = off;   (void) arena_take(&off, sizeof(mfft_move)*S);
   o_slab = off; (void) arena_take(&off, 2*half);
   up = o_slab + half;
   need = off + half + 256;
   if (need > g_arena_bytes)
   {
      mfft_dev_free(g_arena);
      g_arena_bytes = need + need/2;
      g_arena = (unsigned char *) mfft_dev_alloc(g_arena_bytes);
      if (!g_arena) { g_arena_bytes = 0; mfft_die(fn, "device allocation failed: %s", mfft_dev_last_error()); }
   }
   if (up > g_stage_bytes)
   {
      mfft_host_free_pinned(g_stage);
      g_stage_bytes = 2*up;
      g_stage = (limb_t *) mfft_host_alloc_pinned(g_stage_bytes);
      if (!g_stage) { g_stage_bytes = 0; mfft_die(fn, "out of (pinned) host memory"); }
   }
   st = (unsigned char *) g_stage;
   b = (mfft_batch *)(st + o_b); mv = (mfft_move *)(st + o_mv); stage = (limb_t *)(st + o_slab);
   d_b = (mfft_batch *)(g_arena + o_b); d_base = (uint32_t *)(g_arena + o_base);
   ds.s = s; ds.d_ops = (mfft_op *)(g_arena + o_ops); d_mv = (mfft_move *)(g_arena + o_mv);
   d_slab = (limb_t *)(g_arena + o_slab); d_dst = (limb_t *)(g_arena + off);
   b->base = 0; b->parity = 0; b->col = col; b->pad = 0;
   *(uint32_t *)(st + o_base) = 0;
   if (s->nops) memcpy(st + o_ops, s->ops, sizeof(mfft_op)*s->nops);
   for (k = 0; k < S; k++) { mv[k].src_slot = s->slot[k]; mv[k].dst_pos = k; }
   for (k = 0; k < S; k++)
      if (in[k]) memcpy(stage + (size_t) k*pitch, in[k], pitch*sizeof(limb_t));
      else memset(stage + (size_t) k*pitch, 0, pitch*sizeof(limb_t));
   g.S = S; g.slot_stride = 1; g.half_blocks = S; g.l = l; g.pitch = pitch;
   if (mfft_dev_h2d(g_arena, st, up, NULL) ||
       mfft_dsched_run(&ds, d_slab, &g, d_b, 1, NULL) ||
       mfft_dev_finalize(d_dst, 1, d_base, d_slab, &g, d_mv, S, d_b, 1, 0, normalise, NULL) ||
       mfft_dev_d2h(stage, d_dst, half, NULL) || mfft_dev_sync(NULL))
      mfft_die(fn, "device execution failed: %s", mfft_dev_last_error());
   for (k = 0; k < S; k++) if (out[k]) memcpy(out[k], stage + (size_t) k*pitch, pitch*sizeof(limb_t));
   mfft_sched_free(s);
   mfft_unlock();
}

static void check_ring(const char *fn, mp_size_t n, mp_bitcnt_t w)
{
   if (n <= 0 || (n & (n - 1)) || w == 0 || ((uint64_t) n*w) % 64)
      mfft_die(fn, "illegal ring: n=%ld w=%lu (need n a power of two and 64 | n*w)", (long) n, (unsigned long) w);
}

/* all 1-D transforms share this: positions k < 2n are ii[k*is] */
/* where the results go for the entry points that take (rr, rs): result k into the block rr[k*rs] points
   to (the reference stores the result pointers there, mul_fft.c:797-802, 821-826; every caller passes
   rr == ii, rs == 1) */
static mp_limb_t **g_rr = NULL; static mp_size_t g_rs = 1;

static void transform_1d(const char *fn, mfft_transform_kind kind, mp_limb_t **ii, mp_size_t is, mp_size_t n,
                         mp_bitcnt_t w, mp_size_t ws, mp_size_t r, mp_size_t c, mp_size_t rs, mp_size_t trunc)
{
   uint32_t S, k; mfft_sched *s; mp_limb_t **tab, **otab = NULL;
   mp_limb_t **rr = g_rr; const mp_size_t rrs = g_rs;
   g_rr = NULL; g_rs = 1;
   check_ring(fn, n, w);
   if (is <= 0) mfft_die(fn, "illegal stride %ld", (long) is);
   S = (uint32_t)(2*n);
   s = mfft_sched_new(S, (uint64_t) n*w);
   tab = (mp_limb_t **) malloc(sizeof(mp_limb_t *) * S);
   if (!s || !tab) mfft_die(fn, "out of host memory");
   if (mfft_sched_emit(s, kind, 0, 1, (uint64_t) n, w, (uint64_t) ws, (uint64_t) r, (uint64_t) rs, (uint64_t) trunc) != 0)
      mfft_die(fn, "illegal transform parameters (n=%ld w=%lu trunc=%ld; trunc must be even, 2 <= trunc <= 2n)",
               (long) n, (unsigned long) w, (long) trunc);
   for (k = 0; k < S; k++) tab[k] = ii[(size_t) k*is];
   if (rr && (rr != ii || rrs != is))
   {
      if (rrs <= 0) mfft_die(fn, "illegal result stride %ld", (long) rrs);
      otab = (mp_limb_t **) malloc(sizeof(mp_limb_t *) * S);
      if (!otab) mfft_die(fn, "out of host memory");
      for (k = 0; k < S; k++) otab[k] = rr[(size_t) k*rrs];
   }
   run_on_host_blocks(fn, s, (uint32_t)((uint64_t) n*w/64), tab, otab ? otab : tab, (uint32_t) c, 0);
   free(tab); free(otab);
}

void FFT_radix2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("FFT_radix2", MFFT_T_FFT, ii, 1, n, w, 0, 0, 0, 0, 0); mfft_unlock(); }

void IFFT_radix2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                 mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("IFFT_radix2", MFFT_T_IFFT, ii, 1, n, w, 0, 0, 0, 0, 0); mfft_unlock(); }

void FFT_radix2_truncate(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                         mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("FFT_radix2_truncate", MFFT_T_FFT_TRUNC, ii, 1, n, w, 0, 0, 0, 0, trunc); mfft_unlock(); }

void FFT_radix2_truncate1(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                          mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("FFT_radix2_truncate1", MFFT_T_FFT_TRUNC1, ii, 1, n, w, 0, 0, 0, 0, trunc); mfft_unlock(); }

void IFFT_radix2_truncate(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                          mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("IFFT_radix2_truncate", MFFT_T_IFFT_TRUNC, ii, 1, n, w, 0, 0, 0, 0, trunc); mfft_unlock(); }

void IFFT_radix2_truncate1(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                           mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("IFFT_radix2_truncate1", MFFT_T_IFFT_TRUNC1, ii, 1, n, w, 0, 0, 0, 0, trunc); mfft_unlock(); }

void FFT_radix2_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                        mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c, mp_size_t rs)
{ (void) t1; (void) t2; (void) temp;
  transform_1d("FFT_radix2_twiddle", MFFT_T_FFT, ii, is, n, w, ws, r, c, rs, 0); }

void IFFT_radix2_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                         mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c, mp_size_t rs)
{ (void) t1; (void) t2; (void) temp;
  transform_1d("IFFT_radix2_twiddle", MFFT_T_IFFT, ii, is, n, w, ws, r, c, rs, 0); }

void FFT_radix2_truncate_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                                 mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c,
                                 mp_size_t rs, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp;
  transform_1d("FFT_radix2_truncate_twiddle", MFFT_T_FFT_TRUNC, ii, is, n, w, ws, r, c, rs, trunc); }

void FFT_radix2_truncate1_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                                  mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c,
                                  mp_size_t rs, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp;
  transform_1d("FFT_radix2_truncate1_twiddle", MFFT_T_FFT_TRUNC1, ii, is, n, w, ws, r, c, rs, trunc); }

void IFFT_radix2_truncate_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                                  mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c,
                                  mp_size_t rs, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp;
  transform_1d("IFFT_radix2_truncate_twiddle", MFFT_T_IFFT_TRUNC, ii, is, n, w, ws, r, c, rs, trunc); }

void IFFT_radix2_truncate1_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                                   mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c,
                                   mp_size_t rs, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp;
  transform_1d("IFFT_radix2_truncate1_twiddle", MFFT_T_IFFT_TRUNC1, ii, is, n, w, ws, r, c, rs, trunc); }

void FFT_radix2_negacyclic(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                           mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("FFT_radix2_negacyclic", MFFT_T_FFT_NEGACYCLIC, ii, 1, n, w, 0, 0, 0, 0, 0); mfft_unlock(); }

void IFFT_radix2_negacyclic(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                            mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs;
  transform_1d("IFFT_radix2_negacyclic", MFFT_T_IFFT_NEGACYCLIC, ii, 1, n, w, 0, 0, 0, 0, 0); mfft_unlock(); }

/* 1-D transforms of length 4n with the root sqrt2^w (mul_fft.c:839, 1488, 1230, 1792) */
static void transform_sqrt2_1d(const char *fn, int inverse, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_size_t trunc)
{
   uint32_t S, k; mfft_sched *s; mp_limb_t **tab, **otab = NULL;
   mp_limb_t **rr = g_rr; const mp_size_t rrs = g_rs;
   g_rr = NULL; g_rs = 1;
   check_ring(fn, n, w);
   S = (uint32_t)(4*n);
   s = mfft_sched_new(S, (uint64_t) n*w);
   tab = (mp_limb_t **) malloc(sizeof(mp_limb_t *) * S);
   if (!s || !tab) mfft_die(fn, "out of host memory");
   if (mfft_sched_emit_sqrt2_1d(s, inverse, (uint64_t) n, w, (uint64_t) trunc) != 0)
      mfft_die(fn, "illegal transform parameters (n=%ld w=%lu trunc=%ld; trunc must be even, 2n < trunc <= 4n, 4 | n*w)",
               (long) n, (unsigned long) w, (long) trunc);
   for (k = 0; k < S; k++) tab[k] = ii[k];
   if (rr && (rr != ii || rrs != 1))
   {
      if (rrs <= 0) mfft_die(fn, "illegal result stride %ld", (long) rrs);
      otab = (mp_limb_t **) malloc(sizeof(mp_limb_t *) * S);
      if (!otab) mfft_die(fn, "out of host memory");
      for (k = 0; k < S; k++) otab[k] = rr[(size_t) k*rrs];
   }
   run_on_host_blocks(fn, s, (uint32_t)((uint64_t) n*w/64), tab, otab ? otab : tab, 0, 0);
   free(tab); free(otab);
}

void FFT_radix2_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                      mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs; transform_sqrt2_1d("FFT_radix2_sqrt2", 0, ii, n, w, 4*n); mfft_unlock(); }

void IFFT_radix2_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                       mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs; transform_sqrt2_1d("IFFT_radix2_sqrt2", 1, ii, n, w, 4*n); mfft_unlock(); }

void FFT_radix2_truncate_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                               mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs; transform_sqrt2_1d("FFT_radix2_truncate_sqrt2", 0, ii, n, w, trunc); mfft_unlock(); }

void IFFT_radix2_truncate_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                                mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfft_lock(); g_rr = rr; g_rs = rs; transform_sqrt2_1d("IFFT_radix2_truncate_sqrt2", 1, ii, n, w, trunc); mfft_unlock(); }

/* ------------------------------- MFA on host pointer tables -------------------------------- */
static void mfa_host(const char *fn, int inverse, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_size_t n1,
                     mp_size_t trunc, int truncated, int sqrt2)
{
   mfft_mfa m; int rc; uint64_t N, k, i, j; size_t half; limb_t *stage, *d_slab, *d_dst; uint32_t pitch;
   check_ring(fn, n, w);
   mfft_lock();
   mfft_require_device(fn);
   if (sqrt2) rc = mfft_mfa_build_sqrt2(&m, inverse, (uint64_t) n, w, (uint64_t) n1, (uint64_t) trunc, 0, 1);
   else
   rc = mfft_mfa_build(&m, inverse, (uint64_t) n, w, (uint64_t) n1, truncated ? (uint64_t) trunc : 0, 0, truncated);
   if (rc != 0) mfft_die(fn, "illegal MFA parameters n=%ld w=%lu n1=%ld trunc=%ld (code %d; trunc must be a multiple "
                         "of 2*n1, mul_fft.c:2209-2211)", (long) n, (unsigned long) w, (long) n1, (long) trunc, rc);
   N = m.N; pitch = m.pitch; half = (size_t) N*pitch*sizeof(limb_t);
   stage = (limb_t *) calloc((size_t) N*pitch, sizeof(limb_t));
   d_slab = (limb_t *) mfft_dev_alloc(2*half); d_dst = (limb_t *) mfft_dev_alloc(half);
   if (!stage || !d_slab || !d_dst) mfft_die(fn, "allocation failed: %s", mfft_dev_last_error());
   for (k = 0; k < N; k++) memcpy(stage + k*pitch, ii[k], (m.l + 1)*sizeof(limb_t));
   if (mfft_dev_h2d(d_slab, stage, half, NULL) ||
       mfft_dev_h2d(d_dst, stage, half, NULL) ||          /* rows the transform does not produce keep their input */
       mfft_mfa_exec(&m, d_slab, d_dst, NULL) ||
       mfft_dev_d2h(stage, d_dst, half, NULL) || mfft_dev_sync(NULL))
      mfft_die(fn, "device execution failed: %s", mfft_dev_last_error());
   if (!inverse)
   {  /* valid outputs: rows revbin(s), s < trunc_rows, all columns (2392-2408) */
      for (i = 0; i < m.nrows; i++)
         for (j = 0; j < m.n1; j++)
         {  k = (uint64_t) m.rows[i]*m.n1 + j; memcpy(ii[k], stage + k*pitch, (m.l + 1)*sizeof(limb_t)); }
   } else
   {  /* valid outputs: positions j < trunc (2974-2976) */
      for (k = 0; k < m.trunc_rows*m.n1; k++) memcpy(ii[k], stage + k*pitch, (m.l + 1)*sizeof(limb_t));
   }
   mfft_dev_free(d_slab); mfft_dev_free(d_dst); free(stage);
   mfft_mfa_free(&m);
   mfft_unlock();
}

void FFT_radix2_mfa(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                    mp_limb_t **temp, mp_size_t n1)
{ (void) t1; (void) t2; (void) temp; mfa_host("FFT_radix2_mfa", 0, ii, n, w, n1, 0, 0, 0); }

void IFFT_radix2_mfa(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                     mp_limb_t **temp, mp_size_t n1)
{ (void) t1; (void) t2; (void) temp; mfa_host("IFFT_radix2_mfa", 1, ii, n, w, n1, 0, 0, 0); }

void FFT_radix2_mfa_truncate(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                             mp_limb_t **temp, mp_size_t n1, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfa_host("FFT_radix2_mfa_truncate", 0, ii, n, w, n1, trunc, 1, 0); }

void IFFT_radix2_mfa_truncate(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                              mp_limb_t **temp, mp_size_t n1, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfa_host("IFFT_radix2_mfa_truncate", 1, ii, n, w, n1, trunc, 1, 0); }

/* the sqrt2 transforms of length 4n (mul_fft.c:2212, 2593): valid outputs as in the reference -- forward,
   every row of the first half and rows revbin(s), s < (trunc-2n)/n1, of the second; inverse, ii[0..trunc) */
void FFT_radix2_mfa_truncate_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                                   mp_limb_t **temp, mp_size_t n1, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfa_host("FFT_radix2_mfa_truncate_sqrt2", 0, ii, n, w, n1, trunc, 1, 1); }

void IFFT_radix2_mfa_truncate_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                                    mp_limb_t **temp, mp_size_t n1, mp_size_t trunc)
{ (void) t1; (void) t2; (void) temp; mfa_host("IFFT_radix2_mfa_truncate_sqrt2", 1, ii, n, w, n1, trunc, 1, 1); }

/* ----------------------------- single-block primitives ------------------------------------- */
/* positions 0,1 = inputs A,B; outputs S -> position 0, T -> position 1 */
static void two_block_op(const char *fn, uint32_t l, mp_limb_t *outS, mp_limb_t *outT, mp_limb_t *A, mp_limb_t *B,
                         int sSA, uint64_t eSA, int sSB, uint64_t eSB, int sTA, uint64_t eTA, int sTB, uint64_t eTB,
                         int normalise)
{
   mfft_sched *s = mfft_sched_new(2, 64ull*l);
   mp_limb_t *in[2], *out[2];
   if (!s) mfft_die(fn, "out of host memory");
   if (l == 0) mfft_die(fn, "zero-limb ring");
   in[0] = A; in[1] = B; out[0] = outS; out[1] = outT;
   mfft_sched_emit_op(s, 0, B ? 1 : MFFT_NONE, 0, sSA, eSA, sSB, eSB, outT ? 1 : MFFT_NONE, sTA, eTA, sTB, eTB);
   run_on_host_blocks(fn, s, l, in, out, 0, normalise);
}

#define M2OF(l) (128ull*(uint64_t)(l))
#define NEGE(e, l) ((M2OF(l) - ((uint64_t)(e) % M2OF(l))) % M2OF(l))

void mpn_normmod_2expp1(mp_limb_t *t, mp_size_t l)
{ two_block_op("mpn_normmod_2expp1", (uint32_t) l, t, NULL, t, NULL, 1, 0, 0, 0, 0, 0, 0, 0, 1); }

void mpn_mul_2expmod_2expp1(mp_limb_t *t, mp_limb_t *i1, mp_size_t limbs, mp_bitcnt_t d)
{
   if (d >= 64) mfft_die("mpn_mul_2expmod_2expp1", "d=%lu out of range", (unsigned long) d);
   two_block_op("mpn_mul_2expmod_2expp1", (uint32_t) limbs, t, NULL, i1, NULL, 1, d, 0, 0, 0, 0, 0, 0, 0);
}

void mpn_div_2expmod_2expp1(mp_limb_t *t, mp_limb_t *i1, mp_size_t limbs, mp_bitcnt_t d)
{
   if (d >= 64) mfft_die("mpn_div_2expmod_2expp1", "d=%lu out of range", (unsigned long) d);
   two_block_op("mpn_div_2expmod_2expp1", (uint32_t) limbs, t, NULL, i1, NULL, 1, NEGE(d, limbs), 0, 0, 0, 0, 0, 0, 0);
}

void mpn_lshB_sumdiffmod_2expp1(mp_limb_t *t, mp_limb_t *u, mp_limb_t *i1, mp_limb_t *i2, mp_size_t limbs,
                                mp_size_t x, mp_size_t y)
{
   if (x < 0 || y < 0 || x > limbs || y > limbs) mfft_die("mpn_lshB_sumdiffmod_2expp1", "shift out of range");
   two_block_op("mpn_lshB_sumdiffmod_2expp1", (uint32_t) limbs, t, u, i1, i2,
                1, 64ull*x, 1, 64ull*x, 1, 64ull*y, -1, 64ull*y, 0);
}

void mpn_sumdiff_rshBmod_2expp1(mp_limb_t *t, mp_limb_t *u, mp_limb_t *i1, mp_limb_t *i2, mp_size_t limbs,
                                mp_size_t x, mp_size_t y)
{
   if (x < 0 || y < 0 || x > limbs || y > limbs) mfft_die("mpn_sumdiff_rshBmod_2expp1", "shift out of range");
   two_block_op("mpn_sumdiff_rshBmod_2expp1", (uint32_t) limbs, t, u, i1, i2,
                1, NEGE(64ull*x, limbs), 1, NEGE(64ull*y, limbs), 1, NEGE(64ull*x, limbs), -1, NEGE(64ull*y, limbs), 0);
}

void FFT_radix2_butterfly(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t i, mp_size_t n, mp_bitcnt_t w)
{
   uint64_t l, e;
   check_ring("FFT_radix2_butterfly", n, w);
   l = (uint64_t) n*w/64; e = ((uint64_t) i % (2*(uint64_t) n)) * w;
   two_block_op("FFT_radix2_butterfly", (uint32_t) l, s, t, i1, i2, 1, 0, 1, 0, 1, e, -1, e, 0);
}

void FFT_radix2_inverse_butterfly(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t i, mp_size_t n, mp_bitcnt_t w)
{
   uint64_t l, e;
   check_ring("FFT_radix2_inverse_butterfly", n, w);
   l = (uint64_t) n*w/64; e = NEGE(((uint64_t) i % (2*(uint64_t) n)) * w, l);
   two_block_op("FFT_radix2_inverse_butterfly", (uint32_t) l, s, t, i1, i2, 1, 0, 1, e, 1, 0, -1, e, 0);
}

void FFT_radix2_twiddle_butterfly(mp_limb_t *u, mp_limb_t *v, mp_limb_t *s, mp_limb_t *t, mp_size_t NW, mp_bitcnt_t b1, mp_bitcnt_t b2)
{
   uint64_t l = (uint64_t) NW/64;
   if (NW <= 0 || NW % 64) mfft_die("FFT_radix2_twiddle_butterfly", "NW=%ld must be a positive multiple of 64", (long) NW);
   b1 %= 2*(uint64_t) NW; b2 %= 2*(uint64_t) NW;
   two_block_op("FFT_radix2_twiddle_butterfly", (uint32_t) l, u, v, s, t, 1, b1, 1, b1, 1, b2, -1, b2, 0);
}

void FFT_radix2_twiddle_inverse_butterfly(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t NW, mp_bitcnt_t b1, mp_bitcnt_t b2)
{
   uint64_t l = (uint64_t) NW/64;
   if (NW <= 0 || NW % 64) mfft_die("FFT_radix2_twiddle_inverse_butterfly", "NW=%ld must be a positive multiple of 64", (long) NW);
   two_block_op("FFT_radix2_twiddle_inverse_butterfly", (uint32_t) l, s, t, i1, i2,
                1, NEGE(b1, l), 1, NEGE(b2, l), 1, NEGE(b1, l), -1, NEGE(b2, l), 0);
}

void FFT_twiddle(mp_limb_t *r, mp_limb_t *i1, mp_size_t i, mp_size_t n, mp_bitcnt_t w)
{
   uint64_t l, e;
   check_ring("FFT_twiddle", n, w);
   l = (uint64_t) n*w/64; e = ((uint64_t) i % (2*(uint64_t) n)) * w;
   two_block_op("FFT_twiddle", (uint32_t) l, r, NULL, i1, NULL, 1, e, 0, 0, 0, 0, 0, 0, 0);
}

/* mul_fft.c:2078 / 2461: the untruncated pair is the truncated one at trunc = 4n */
void FFT_radix2_mfa_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                          mp_limb_t **temp, mp_size_t n1)
{ (void) t1; (void) t2; (void) temp; mfa_host("FFT_radix2_mfa_sqrt2", 0, ii, n, w, n1, 4*n, 1, 1); }

void IFFT_radix2_mfa_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                           mp_limb_t **temp, mp_size_t n1)
{ (void) t1; (void) t2; (void) temp; mfa_host("IFFT_radix2_mfa_sqrt2", 1, ii, n, w, n1, 4*n, 1, 1); }

/* ---- sqrt2 butterflies (mul_fft.c:591, 673, 972): z1^i = 2^e (2^(nw/2) - 1), e = (i w - 1)/2 + nw/4 ---- */
static void sqrt2_check(const char *fn, mp_size_t i, mp_size_t n, mp_bitcnt_t w)
{
   check_ring(fn, n, w);
   if (!(i & 1) || !(w & 1) || i < 0 || i >= 4*n || ((uint64_t) n*w) % 256)
      mfft_die(fn, "needs odd i < 4n, odd w and 256 | n*w (i=%ld n=%ld w=%lu)", (long) i, (long) n, (unsigned long) w);
}

void FFT_radix2_butterfly_sqrt2(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t i,
                                mp_size_t n, mp_bitcnt_t w, mp_limb_t *temp)
{
   uint64_t NW, e; mfft_sched *sc; mp_limb_t *in[2], *out[2];
   (void) temp;
   sqrt2_check("FFT_radix2_butterfly_sqrt2", i, n, w);
   NW = (uint64_t) n*w; e = (((uint64_t) i*w - 1)/2 + NW/4) % (2*NW);
   if (!(sc = mfft_sched_new(2, NW))) mfft_die("FFT_radix2_butterfly_sqrt2", "out of host memory");
   in[0] = i1; in[1] = i2; out[0] = s; out[1] = t;
   mfft_sched_emit_op(sc, 0, 1, 0, 1, 0, 1, 0, 1, 1, e, -1, e);                     /* a + b, 2^e (a - b) */
   mfft_sched_emit_op(sc, 1, 1, 1, 1, NW/2, -1, 0, MFFT_NONE, 0, 0, 0, 0);          /* x 2^(nw/2) - x */
   run_on_host_blocks("FFT_radix2_butterfly_sqrt2", sc, (uint32_t)(NW/64), in, out, 0, 0);
}

void FFT_radix2_inverse_butterfly_sqrt2(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t i,
                                        mp_size_t n, mp_bitcnt_t w, mp_limb_t *temp)
{
   uint64_t NW, e; mfft_sched *sc; mp_limb_t *in[2], *out[2];
   (void) temp;
   sqrt2_check("FFT_radix2_inverse_butterfly_sqrt2", i, n, w);
   NW = (uint64_t) n*w; e = (4*NW - ((uint64_t) i*w - 1)/2 - 1 + NW/4) % (2*NW);     /* mul_fft.c:653-672 */
   if (!(sc = mfft_sched_new(2, NW))) mfft_die("FFT_radix2_inverse_butterfly_sqrt2", "out of host memory");
   in[0] = i1; in[1] = i2; out[0] = s; out[1] = t;
   mfft_sched_emit_op(sc, 1, MFFT_NONE, 1, 1, e, 0, 0, MFFT_NONE, 0, 0, 0, 0);
   mfft_sched_emit_op(sc, 1, 1, 1, 1, NW/2, -1, 0, MFFT_NONE, 0, 0, 0, 0);
   mfft_sched_emit_op(sc, 0, 1, 0, 1, 0, 1, 0, 1, 1, 0, -1, 0);
   run_on_host_blocks("FFT_radix2_inverse_butterfly_sqrt2", sc, (uint32_t)(NW/64), in, out, 0, 0);
}

void FFT_twiddle_sqrt2(mp_limb_t *r, mp_limb_t *i1, mp_size_t i, mp_size_t n, mp_bitcnt_t w, mp_limb_t *temp)
{
   uint64_t NW, e; mfft_sched *sc; mp_limb_t *in[1], *out[1];
   (void) temp;
   sqrt2_check("FFT_twiddle_sqrt2", i, n, w);
   NW = (uint64_t) n*w; e = (((uint64_t) i*w - 1)/2 + NW/4) % (2*NW);
   if (!(sc = mfft_sched_new(1, NW))) mfft_die("FFT_twiddle_sqrt2", "out of host memory");
   in[0] = i1; out[0] = r;
   mfft_sched_emit_op(sc, 0, MFFT_NONE, 0, 1, e, 0, 0, MFFT_NONE, 0, 0, 0, 0);
   mfft_sched_emit_op(sc, 0, 0, 0, 1, NW/2, -1, 0, MFFT_NONE, 0, 0, 0, 0);
   run_on_host_blocks("FFT_twiddle_sqrt2", sc, (uint32_t)(NW/64), in, out, 0, 0);
}

/* ---------------------------------- split / combine ---------------------------------------- */
mp_size_t FFT_split_bits(mp_limb_t **poly, mp_limb_t *limbs, mp_size_t total_limbs, mp_size_t bits, mp_size_t output_limbs)
{
   uint64_t length, k; uint32_t l = (uint32_t) output_limbs, pitch = l + 1;
   limb_t *d_src, *d_slab, *stage;
   if (total_limbs <= 0 || bits <= 0 || output_limbs <= 0 || (uint64_t) bits > 64ull*(uint64_t) output_limbs)
      mfft_die("FFT_split_bits", "illegal sizes total=%ld bits=%ld output_limbs=%ld", (long) total_limbs, (long) bits, (long) output_limbs);
   length = (64ull*(uint64_t) total_limbs - 1)/(uint64_t) bits + 1;             /* mul_fft.c:118 */
   mfft_lock();
   mfft_require_device("FFT_split_bits");
   d_src = (limb_t *) mfft_dev_alloc((size_t) total_limbs*8);
   d_slab = (limb_t *) mfft_dev_alloc((size_t) length*pitch*8);
   stage = (limb_t *) malloc((size_t) length*pitch*8);
   if (!d_src || !d_slab || !stage) mfft_die("FFT_split_bits", "allocation failed: %s", mfft_dev_last_error());
   if (mfft_dev_h2d(d_src, limbs, (size_t) total_limbs*8, NULL) ||
       mfft_dev_split(d_slab, l, pitch, d_src, (uint64_t) total_limbs, (uint64_t) bits, length, length, NULL) ||
       mfft_dev_d2h(stage, d_slab, (size_t) length*pitch*8, NULL) || mfft_dev_sync(NULL))
      mfft_die("FFT_split_bits", "device execution failed: %s", mfft_dev_last_error());
   for (k = 0; k < length; k++) memcpy(poly[k], stage + k*pitch, pitch*8);
   mfft_dev_free(d_src); mfft_dev_free(d_slab); free(stage);
   mfft_unlock();
   return (mp_size_t) length;
}

mp_size_t FFT_split(mp_limb_t **poly, mp_limb_t *limbs, mp_size_t total_limbs, mp_size_t coeff_limbs, mp_size_t output_limbs)
{ return FFT_split_bits(poly, limbs, total_limbs, 64*coeff_limbs, output_limbs); }

/* Blocks are read as output_limbs+1 limbs (the reference adds the carry limb too when the
 * coefficient is bit-shifted, mul_fft.c:231-232).  Like the reference (185, 229-233) the sum is ADDED to
 * what res holds on entry (which its contract wants zeroed, 203-204), modulo 2^(64 total_limbs). */
void FFT_combine_bits(mp_limb_t *res, mp_limb_t **poly, mp_size_t length, mp_size_t bits, mp_size_t output_limbs, mp_size_t total_limbs)
{
   uint32_t l = (uint32_t) output_limbs + 1, pitch = l; uint64_t k;
   limb_t *d_res, *d_slab, *stage, *sum; void *work; unsigned cy = 0;
   if (length < 0 || bits <= 0 || output_limbs <= 0 || total_limbs <= 0)
      mfft_die("FFT_combine_bits", "illegal sizes length=%ld bits=%ld output_limbs=%ld total=%ld", (long) length, (long) bits, (long) output_limbs, (long) total_limbs);
   mfft_lock();
   mfft_require_device("FFT_combine_bits");
   d_res = (limb_t *) mfft_dev_alloc((size_t) total_limbs*8);
   d_slab = (limb_t *) mfft_dev_alloc((size_t)(length ? length : 1)*pitch*8);
   work = mfft_dev_alloc(mfft_dev_combine_work((uint64_t) total_limbs));
   stage = (limb_t *) malloc((size_t)(length ? length : 1)*pitch*8);
   sum = (limb_t *) malloc((size_t) total_limbs*8);
   if (!d_res || !d_slab || !work || !stage || !sum) mfft_die("FFT_combine_bits", "allocation failed: %s", mfft_dev_last_error());
   for (k = 0; k < (uint64_t) length; k++) memcpy(stage + k*pitch, poly[k], pitch*8);
   if (mfft_dev_h2d(d_slab, stage, (size_t) length*pitch*8, NULL) ||
       mfft_dev_combine(d_res, (uint64_t) total_limbs, d_slab, l, pitch, (uint64_t) bits, (uint64_t) length, work, NULL) ||
       mfft_dev_d2h(sum, d_res, (size_t) total_limbs*8, NULL) || mfft_dev_sync(NULL))
      mfft_die("FFT_combine_bits", "device execution failed: %s", mfft_dev_last_error());
   for (k = 0; k < (uint64_t) total_limbs; k++)
   {
      const limb_t a = res[k], t = a + sum[k], u = t + cy;
      cy = (unsigned)((t < a) | (u < t));
      res[k] = u;
   }
   mfft_dev_free(d_res); mfft_dev_free(d_slab); mfft_dev_free(work); free(stage); free(sum);
   mfft_unlock();
}

void FFT_combine(mp_limb_t *res, mp_limb_t **poly, mp_size_t length, mp_size_t coeff_limbs, mp_size_t output_limbs, mp_size_t total_limbs)
{
   /* the limb-aligned variant adds exactly output_limbs limbs per coefficient (mul_fft.c:188);
      clear the carry limbs of a private copy so that both variants share one kernel */
   mp_size_t k; mp_limb_t **tab, *copy; size_t sz = (size_t) output_limbs + 1;
   if (length <= 0) return;
   tab = (mp_limb_t **) malloc(sizeof(mp_limb_t *) * (size_t) length);
   copy = (mp_limb_t *) malloc(sz*8*(size_t) length);
   if (!tab || !copy) mfft_die("FFT_combine", "out of host memory");
   for (k = 0; k < length; k++)
   {
      tab[k] = copy + (size_t) k*sz; memcpy(tab[k], poly[k], (sz - 1)*8); tab[k][sz - 1] = 0;
   }
   FFT_combine_bits(res, tab, length, 64*coeff_limbs, output_limbs, total_limbs);
   free(tab); free(copy);
}

/* ---------------------------------- mulmod 2^bits + 1 -------------------------------------- */
/* plans of the transform-based path, keyed by (limbs, count) */
#define MM_CACHE 4
static mpirfft_mulmod_plan *g_mm[MM_CACHE]; static size_t g_mm_l[MM_CACHE], g_mm_count[MM_CACHE]; static unsigned g_mm_next = 0;

int mpirfft_mulmod_batch_device(mp_limb_t *d_a, const mp_limb_t *d_b, size_t count, size_t l, size_t pitch, void *stream)
{
   uint32_t *idx, *d_idx; size_t k; int rc;
   mp_bitcnt_t depth, w;
   if (!count) return 0;
   if (l == 0 || pitch < l + 1 || count > 0xffffffffu) return MPIRFFT_EINVAL;
   if (l >= 250 && l != 256 && l != 512 && mpirfft_mulmod_params((mp_size_t) l, &depth, &w) == 0)
   {  /* large residues: negacyclic transform over a 64..512-limb ring, like the reference's
         fft_mulmod_2expp1 from 250 limbs up (mul_fft.c:3135-3164); 256 and 512 limbs are served
         directly by the warp-level product kernel */
      mpirfft_mulmod_plan *pl = NULL; int i;
      mfft_lock();                 /* held across cache lookup, eviction and execution */
      for (i = 0; i < MM_CACHE; i++) if (g_mm[i] && g_mm_l[i] == l && g_mm_count[i] == count) pl = g_mm[i];
      if (!pl)
      {
         if ((rc = mpirfft_mulmod_plan_create(&pl, (mp_size_t) l, depth, w, count)) != 0) { mfft_unlock(); return rc; }
         i = (int)(g_mm_next++ % MM_CACHE);
         if (g_mm[i]) mpirfft_mulmod_plan_destroy(g_mm[i]);
         g_mm[i] = pl; g_mm_l[i] = l; g_mm_count[i] = count;
      }
      rc = mpirfft_mulmod_plan_exec(pl, d_a, d_a, d_b, pitch, -1, stream);
      if (rc == 0 && mfft_dev_sync(stream)) rc = MPIRFFT_ENODEV;
      mfft_unlock();
      return rc;
   }
   mfft_lock();
   if ((rc = mfft_try_device()) != 0) { mfft_unlock(); return rc; }
   idx = (uint32_t *) malloc(sizeof(uint32_t)*count);
   if (!idx) { mfft_unlock(); return MPIRFFT_ENOMEM; }
   for (k = 0; k < count; k++) idx[k] = (uint32_t) k;
   d_idx = (uint32_t *) mfft_upload(idx, sizeof(uint32_t)*count);
   free(idx);
   if (!d_idx) { mfft_unlock(); return MPIRFFT_ENODEV; }
   rc = mfft_dev_pointwise((limb_t *) d_a, (const limb_t *) d_b, d_idx, (uint32_t) count, (uint32_t) l, (uint32_t) pitch, stream);
   if (rc == 0) rc = mfft_dev_sync(stream);
   mfft_dev_free(d_idx);
   mfft_unlock();
   return rc ? MPIRFFT_ENODEV : 0;
}

/* one product on host blocks: a = {i1, top ta}, b = {i2, top tb}; returns the top bit */
static mp_limb_t mulmod_host(const char *fn, mp_limb_t *r, const mp_limb_t *i1, mp_limb_t ta, const mp_limb_t *i2,
                             mp_limb_t tb, uint32_t l)
{
   size_t bytes = ((size_t) l + 1)*8; limb_t *ha, *hb, *da, *db; mp_limb_t top; int rc;
   ha = (limb_t *) malloc(bytes); hb = (limb_t *) malloc(bytes);
   if (!ha || !hb) mfft_die(fn, "out of host memory");
   memcpy(ha, i1, (size_t) l*8); ha[l] = ta; memcpy(hb, i2, (size_t) l*8); hb[l] = tb;
   mfft_lock(); mfft_require_device(fn);
   da = (limb_t *) mfft_upload(ha, bytes); db = (limb_t *) mfft_upload(hb, bytes);
   mfft_unlock();
   if (!da || !db) mfft_die(fn, "device allocation failed: %s", mfft_dev_last_error());
   rc = mpirfft_mulmod_batch_device((mp_limb_t *) da, (const mp_limb_t *) db, 1, l, (size_t) l + 1, NULL);
   if (rc != 0 || mfft_dev_d2h(ha, da, bytes, NULL) || mfft_dev_sync(NULL))
      mfft_die(fn, "device execution failed: %s", mfft_dev_last_error());
   memcpy(r, ha, (size_t) l*8); top = ha[l];
   mfft_dev_free(da); mfft_dev_free(db); free(ha); free(hb);
   return top;
}

mp_limb_t new_mpn_mulmod_2expp1(mp_limb_t *r, mp_limb_t *i1, mp_limb_t *i2, mp_limb_t c, mp_limb_t bits, mp_limb_t *tt)
{
   (void) tt;
   if (bits == 0 || bits % 64) mfft_die("new_mpn_mulmod_2expp1", "bits=%lu must be a positive multiple of 64", (unsigned long) bits);
   return mulmod_host("new_mpn_mulmod_2expp1", r, i1, c & 1, i2, (c >> 1) & 1, (uint32_t)(bits/64));
}

mp_limb_t fft_mulmod_2expp1(mp_limb_t *r, mp_limb_t *i1, mp_limb_t *i2, mp_size_t n, mp_size_t w, mp_limb_t *tt)
{
   uint64_t bits = (uint64_t) n*(uint64_t) w; uint32_t l; mp_limb_t top;
   (void) tt;
   if (n <= 0 || w <= 0 || bits % 64) mfft_die("fft_mulmod_2expp1", "n*w=%lu must be a positive multiple of 64", (unsigned long) bits);
   l = (uint32_t)(bits/64);
   top = mulmod_host("fft_mulmod_2expp1", r, i1, i1[l], i2, i2[l], l);
   if (l >= 250) r[l] = top;       /* the reference's large path writes limbs+1 limbs (3164, 3083) */
   return top;
}

void FFT_mulmod_2expp1(mp_limb_t *r1, mp_limb_t *i1, mp_limb_t *i2, mp_size_t r_limbs, mp_bitcnt_t depth, mp_bitcnt_t w)
{
   (void) depth; (void) w;
   if (r_limbs <= 0) mfft_die("FFT_mulmod_2expp1", "r_limbs=%ld", (long) r_limbs);
   r1[r_limbs] = mulmod_host("FFT_mulmod_2expp1", r1, i1, 0, i2, 0, (uint32_t) r_limbs);
}
