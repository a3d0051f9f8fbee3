/* mul.c -- the integer multiplication driver: new_mpn_mul (mul_fft.c:3190-3265) on device slabs.
 *
 * Pipeline (every step a kernel or a stage list; the slabs never leave HBM):
 *   split i1 -> X.half0      FFT_split_bits + zero fill            3234-3236
 *   fwd MFA  X -> Z.half0    FFT_radix2_mfa_truncate               3237
 *   split i2 -> X.half0 ; fwd MFA X -> Y                           3239-3242
 *   pointwise Z.half0 *= Y on the valid rows                        3244-3253 (row set corrected)
 *   inv MFA  Z -> X.half0, scaled by 2^-(depth+1), normalised       3255-3260
 *   combine  X.half0 -> r                                           3261-3262
 */
#define _POSIX_C_SOURCE 200809L
#include "runtime.h"
#include "hostcopy.h"
#include <sched.h>
#include "../../../include/mpirfft_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MUL_STAGE_EVENTS 32          /* chunks of the result copy */
#define MUL_STAGE_MAX ((size_t) 1 << 30)   /* larger products copy straight from / to the caller's buffers */
struct mpirfft_mul_plan {
   mp_size_t n1, n2; mp_bitcnt_t depth, w;
   int sqrt2;                  /* new_mpn_mul6: transform length 4n with the sqrt2 trick (mul_fft.c:3573) */
   mpirfft_mul_params p;
   uint32_t l, pitch;
   mfft_mfa fwd, inv;
   limb_t *X, *Z, *Y;          /* 2N, 2N, N blocks */
   limb_t *X2;                 /* work slab of the second forward transform (runs concurrently with the first) */
   unsigned char *h_stage;     /* pinned staging for operands / results in pageable host memory (hostcopy.c) */
   size_t h_stage_bytes;
   void *ev_chunk[MUL_STAGE_EVENTS];
   mpirfft_smul_plan *big;     /* coefficient rings above 512 limbs: the sharded plan on one rank (multi-layer sliced
                                  passes, pointwise products through the negacyclic recursion) with local copies for
                                  the exchanges */
   void *s_fwd2, *ev_fork, *ev_join;
   uint32_t *d_pw_blocks; uint32_t npw;
   void *combine_work;
   limb_t *d_i1, *d_i2, *d_r;  /* staging for the host-pointer entry */
   void *s_copy, *s_comp, *ev1, *ev2;   /* copy / compute streams of the host-pointer entry */
   size_t dev_bytes;
};

int mpirfft_mul_params_get(mpirfft_mul_params *o, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth, mp_bitcnt_t w)
{
   uint64_t n, nw;
   memset(o, 0, sizeof(*o));
   if (n1 <= 0 || n2 <= 0 || depth < 2 || depth > 26 || w == 0 || w > 4096) return MPIRFFT_EINVAL;
   n = (uint64_t)1 << depth; nw = n*w;
   if (nw % 64) return MPIRFFT_EINVAL;                       /* mul_fft.c:48 */
   if (nw <= depth + 1) return MPIRFFT_EINVAL;
   o->n = n;
   o->bits1 = (nw - depth)/2;                                /* 3194 */
   o->sqrt = (uint64_t)1 << (depth/2);                       /* 3195 */
   o->j1 = ((uint64_t) n1*64 - 1)/o->bits1 + 1;              /* 3198 */
   o->j2 = ((uint64_t) n2*64 - 1)/o->bits1 + 1;              /* 3199 */
   o->trunc = ((o->j1 + o->j2 - 2 + 2*o->sqrt)/(2*o->sqrt))*2*o->sqrt;   /* 3200 */
   o->limbs = nw/64;                                         /* 3202 */
   o->n2 = 2*n/o->sqrt;
   o->trunc_rows = o->trunc/o->sqrt;
   if (o->j1 + o->j2 - 1 > 2*n || o->trunc > 2*n) return MPIRFFT_EINVAL;   /* would segfault, 3186 */
   if (o->n2 < 4 || o->sqrt < 2) return MPIRFFT_EINVAL;
   return 0;
}

/* new_mpn_mul6 (mul_fft.c:3573-3668): transform length 4n, one bit less per coefficient (3578), the
   first 2n coefficients always live, trunc in (2n, 4n] */
int mpirfft_mul6_params_get(mpirfft_mul_params *o, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth, mp_bitcnt_t w)
{
   uint64_t n, nw;
   memset(o, 0, sizeof(*o));
   if (n1 <= 0 || n2 <= 0 || depth < 3 || depth > 25 || w == 0 || w > 4096) return MPIRFFT_EINVAL;
   n = (uint64_t)1 << depth; nw = n*w;
   if (nw % 64 || nw % 4) return MPIRFFT_EINVAL;
   if (nw <= depth + 2) return MPIRFFT_EINVAL;
   o->n = n;
   o->bits1 = (nw - (depth + 1))/2;                          /* 3578 */
   o->sqrt = (uint64_t)1 << (depth/2);                       /* 3577 */
   o->j1 = ((uint64_t) n1*64 - 1)/o->bits1 + 1;              /* 3581 */
   o->j2 = ((uint64_t) n2*64 - 1)/o->bits1 + 1;              /* 3582 */
   o->trunc = 2*o->sqrt*((o->j1 + o->j2 + 2*o->sqrt - 2)/(2*o->sqrt));   /* 3612 */
   o->limbs = nw/64;
   o->n2 = 2*n/o->sqrt;
   if (o->j1 + o->j2 - 1 > 4*n || o->trunc > 4*n) return MPIRFFT_EINVAL;
   if (o->trunc <= 2*n) return MPIRFFT_EINVAL;   /* the second half would be empty (the reference's truncate1 then recurses forever): use new_mpn_mul */
   if (o->n2 < 4 || o->sqrt < 2) return MPIRFFT_EINVAL;
   o->trunc_rows = o->trunc/o->sqrt;                         /* n2 first-half rows + the live second-half rows */
   return 0;
}

/* sqrt2 != NULL: the new_mpn_mul6 shape (transform length 4n) is a candidate too; *sqrt2 says which */
static int choose(mp_size_t n1, mp_size_t n2, mp_bitcnt_t *depth, mp_bitcnt_t *w, int *sqrt2)
{
   mp_bitcnt_t d, ww;
   mpirfft_mul_params p;
   if (sqrt2) *sqrt2 = 0;
   /* first choice: the smallest coefficient ring the fused tile executor and the warp-level product
      kernel handle (64, 128, 256, 512 limbs) -- the pointwise work grows with the ring, the transform
      work shrinks; measured on B200 the smaller ring wins whenever it is legal.  With the sqrt2 trick a
      ring carries twice as many coefficients, so a product that would need the next ring size stays
      one size smaller (a quarter of the multiply work per coefficient, twice the coefficients). */
   {
      static const uint32_t ring[4] = { 64, 128, 256, 512 };
      int i, sq;
      for (i = 0; i < 4; i++)
         for (sq = 0; sq <= (sqrt2 ? 1 : 0); sq++)
            for (ww = 2; ww >= 1; ww--)                  /* w = 2: one layer less for the same ring */
            {
               uint64_t nn = 64ull*ring[i]/ww;
               if (nn & (nn - 1)) continue;
               for (d = 0; ((uint64_t)1 << d) < nn; d++) ;
               if ((sq ? mpirfft_mul6_params_get(&p, n1, n2, d, ww) : mpirfft_mul_params_get(&p, n1, n2, d, ww)) == 0)
               { *depth = d; *w = ww; if (sqrt2) *sqrt2 = sq; return 0; }
            }
   }
   /* smallest coefficient size first: n*w ascending, preferring w = 1 */
   for (d = 6; d <= 26; d++)
      for (ww = 1; ww <= 2; ww++)
         if (mpirfft_mul_params_get(&p, n1, n2, d, ww) == 0)
         {
            /* a deeper transform with w=1 has the same ring as this depth with w=2; take the first hit */
            *depth = d; *w = ww; return 0;
         }
   return MPIRFFT_EINVAL;
}

int mpirfft_choose_params(mp_size_t n1, mp_size_t n2, mp_bitcnt_t *depth, mp_bitcnt_t *w)
{ return choose(n1, n2, depth, w, NULL); }

int mpirfft_choose_params6(mp_size_t n1, mp_size_t n2, mp_bitcnt_t *depth, mp_bitcnt_t *w, int *sqrt2)
{ return choose(n1, n2, depth, w, sqrt2); }

void mpirfft_mul_plan_destroy(mpirfft_mul_plan *pl)
{
   if (!pl) return;
   if (pl->big) mpirfft_smul_plan_destroy(pl->big);
   mfft_lock();
   mfft_mfa_free(&pl->fwd); mfft_mfa_free(&pl->inv);
   mfft_dev_free(pl->X); mfft_dev_free(pl->Z); mfft_dev_free(pl->Y); mfft_dev_free(pl->X2);
   mfft_dev_stream_destroy(pl->s_fwd2); mfft_dev_event_destroy(pl->ev_fork); mfft_dev_event_destroy(pl->ev_join);
   mfft_dev_free(pl->d_pw_blocks); mfft_dev_free(pl->combine_work);
   mfft_dev_free(pl->d_i1); mfft_dev_free(pl->d_i2); mfft_dev_free(pl->d_r);
   mfft_host_free_pinned(pl->h_stage);
   { int k; for (k = 0; k < MUL_STAGE_EVENTS; k++) mfft_dev_event_destroy(pl->ev_chunk[k]); }
   mfft_dev_stream_destroy(pl->s_copy); mfft_dev_stream_destroy(pl->s_comp);
   mfft_dev_event_destroy(pl->ev1); mfft_dev_event_destroy(pl->ev2);
   mfft_unlock();
   free(pl);
}

static int plan_create(mpirfft_mul_plan **out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth, mp_bitcnt_t w, int sqrt2);

int mpirfft_mul_plan_create(mpirfft_mul_plan **out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth, mp_bitcnt_t w)
{ return plan_create(out, n1, n2, depth, w, 0); }

int mpirfft_mul6_plan_create(mpirfft_mul_plan **out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth, mp_bitcnt_t w)
{ return plan_create(out, n1, n2, depth, w, 1); }

static int plan_create(mpirfft_mul_plan **out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth, mp_bitcnt_t w, int sqrt2)
{
   mpirfft_mul_plan *pl; int rc; uint64_t N, i, j; size_t half; uint32_t *blocks = NULL;
   *out = NULL;
   pl = (mpirfft_mul_plan *) calloc(1, sizeof(*pl));
   if (!pl) return MPIRFFT_ENOMEM;
   pl->n1 = n1; pl->n2 = n2; pl->depth = depth; pl->w = w; pl->sqrt2 = sqrt2;
   rc = sqrt2 ? mpirfft_mul6_params_get(&pl->p, n1, n2, depth, w) : mpirfft_mul_params_get(&pl->p, n1, n2, depth, w);
   if (rc != 0) { free(pl); return rc; }
   if (!sqrt2 && pl->p.limbs > 512 && pl->p.limbs % 64 == 0)
   {  /* no shared-memory tile holds such a coefficient: run the plan of the sharded multiplication with
         world size 1 (smul.c) -- its transforms work in place on chunk slices, its pointwise products
         recurse (fft_mulmod_2expp1, mul_fft.c:3125) */
      const char *e = getenv("MPIRFFT_BIG_PLAN");
      if (!(e && e[0] == '0') && mpirfft_smul_plan_create(&pl->big, n1, n2, depth, w, 0, 1) == 0)
      {
         pl->l = (uint32_t) pl->p.limbs; pl->pitch = mfft_pitch(pl->l);
         *out = pl;
         return 0;
      }
      pl->big = NULL;
   }
   mfft_lock();
   if ((rc = mfft_try_device()) != 0) goto fail;
   pl->l = (uint32_t) pl->p.limbs; pl->pitch = mfft_pitch(pl->l);
   N = (sqrt2 ? 4 : 2)*pl->p.n;
   if (sqrt2)
   {  /* FFT/IFFT_radix2_mfa_truncate_sqrt2 (3619, 3625, 3656); / 2^(depth+2) and the normalisation
         (3659-3663) folded into the last inverse pass */
      mfft_mfa_promise_zero_inputs(pl->p.j1 > pl->p.j2 ? pl->p.j1 : pl->p.j2);     /* one plan serves both operands */
      if ((rc = mfft_mfa_build_sqrt2(&pl->fwd, 0, pl->p.n, w, pl->p.sqrt, pl->p.trunc, 0, 1)) != 0) goto fail;
      if ((rc = mfft_mfa_build_sqrt2(&pl->inv, 1, pl->p.n, w, pl->p.sqrt, pl->p.trunc,
                                     (uint32_t)(128ull*pl->l - (depth + 2)), 1)) != 0) goto fail;
   } else
   {
   mfft_mfa_promise_zero_inputs(pl->p.j1 > pl->p.j2 ? pl->p.j1 : pl->p.j2);        /* one plan serves both operands */
   if ((rc = mfft_mfa_build(&pl->fwd, 0, pl->p.n, w, pl->p.sqrt, pl->p.trunc, 0, 1)) != 0) goto fail;
   /* the inverse is unscaled: fold / 2^(depth+1) and the normalisation into its last pass (3256-3260) */
   if ((rc = mfft_mfa_build(&pl->inv, 1, pl->p.n, w, pl->p.sqrt, pl->p.trunc,
                            (uint32_t)(128ull*pl->l - (depth + 1)), 1)) != 0) goto fail;
   }
   half = (size_t) N * pl->pitch * sizeof(limb_t);
   rc = MPIRFFT_ENOMEM;
   pl->X = (limb_t *) mfft_dev_alloc(2*half);
   pl->Z = (limb_t *) mfft_dev_alloc(2*half);
   pl->Y = (limb_t *) mfft_dev_alloc(half);
   pl->combine_work = mfft_dev_alloc(mfft_dev_combine_work((uint64_t)(n1 + n2)));
   if (!pl->X || !pl->Z || !pl->Y || !pl->combine_work) goto fail;
   pl->dev_bytes = 5*half + mfft_dev_combine_work((uint64_t)(n1 + n2));
   {  /* The two forward transforms are independent: on two streams their passes interleave on the SMs,
         one transform's tiles computing while the other's are being fetched (within one launch all
         CTAs move through load -> stages -> store in lockstep waves).  MPIRFFT_FWD_STREAMS=1 serialises. */
      const char *e = getenv("MPIRFFT_FWD_STREAMS");
      if (!(e && e[0] == '1'))
      {
         pl->X2 = (limb_t *) mfft_dev_alloc(2*half);
         pl->s_fwd2 = mfft_dev_stream_create(); pl->ev_fork = mfft_dev_event_create(); pl->ev_join = mfft_dev_event_create();
         if (!pl->X2 || !pl->s_fwd2 || !pl->ev_fork || !pl->ev_join) goto fail;
         pl->dev_bytes += 2*half;
      }
   }
   /* the rows whose coefficients get multiplied: revbin(s, depth+1-depth/2), s < trunc/sqrt
      (mul_fft.c:3244-3253 with the bit width of 3246 corrected, cf. 3629, 3642); for new_mpn_mul6 all
      rows of the first half and those rows of the second (3630-3653); fwd.rows lists them either way */
   pl->npw = (uint32_t)(pl->p.trunc_rows * pl->p.sqrt);
   blocks = (uint32_t *) malloc(sizeof(uint32_t) * pl->npw);
   if (!blocks) goto fail;
   for (i = 0; i < pl->p.trunc_rows; i++)
      for (j = 0; j < pl->p.sqrt; j++)
         blocks[i*pl->p.sqrt + j] = (uint32_t)(pl->fwd.rows[i]*pl->p.sqrt + j);
   pl->d_pw_blocks = (uint32_t *) mfft_upload(blocks, sizeof(uint32_t) * pl->npw);
   free(blocks);
   if (!pl->d_pw_blocks) { rc = MPIRFFT_ENODEV; goto fail; }
   mfft_unlock();
   *out = pl;
   return 0;
fail:
   mfft_unlock();
   mpirfft_mul_plan_destroy(pl);
   return rc;
}

size_t mpirfft_mul_plan_device_bytes(const mpirfft_mul_plan *pl) { return pl->dev_bytes; }   /* (0 for the big-ring plan: its buffers belong to the smul plan) */

uint64_t mpirfft_mul_plan_launches(const mpirfft_mul_plan *pl)
{
   if (pl->big) return 0;     /* not tabulated for the big-ring plan */
   /* 2 x (split + fwd) + pointwise + inv + combine (4 kernels) */
   return 2*((mfft_mfa_can_fuse_split(&pl->fwd) ? 0 : 1) + mfft_mfa_launches(&pl->fwd)) + 1 + mfft_mfa_launches(&pl->inv) + 4;
}

static int exec_phase(mpirfft_mul_plan *pl, int phase, mp_limb_t *d_r, const mp_limb_t *d_i1,
                      const mp_limb_t *d_i2, void *stream);

/* Thread contract: the launches of a phase are issued under the library lock with the library's
 * device made current in the calling thread, so plans may be driven from any host thread.  One
 * plan owns one set of slabs: do not run the same plan concurrently on two streams. */
int mpirfft_mul_exec_phase(mpirfft_mul_plan *pl, int phase, mp_limb_t *d_r, const mp_limb_t *d_i1,
                           const mp_limb_t *d_i2, void *stream)
{
   int rc;
   if (pl->big) return MPIRFFT_EINVAL;        /* the big-ring plan runs as a whole (mpirfft_mul_exec_device) */
   mfft_lock();
   rc = mfft_try_device();
   if (rc == 0) rc = exec_phase(pl, phase, d_r, d_i1, d_i2, stream);
   mfft_unlock();
   return rc;
}

static int exec_phase(mpirfft_mul_plan *pl, int phase, mp_limb_t *d_r, const mp_limb_t *d_i1,
                      const mp_limb_t *d_i2, void *stream)
{
   const mpirfft_mul_params *p = &pl->p;
   int rc = 0;
   switch (phase)
   {
   case 0:
      if (mfft_mfa_can_fuse_split(&pl->fwd))
      { rc = mfft_mfa_exec_split(&pl->fwd, pl->X, pl->Z, (const limb_t *) d_i1, (uint64_t) pl->n1, p->bits1, p->j1, stream); break; }
      if (mfft_dev_split(pl->X, pl->l, pl->pitch, (const limb_t *) d_i1, (uint64_t) pl->n1, p->bits1, p->j1, p->trunc, stream)) return MPIRFFT_ENODEV;
      rc = mfft_mfa_exec(&pl->fwd, pl->X, pl->Z, stream);
      break;
   case 1:
   {
      limb_t *W = pl->X2 ? pl->X2 : pl->X;
      if (mfft_mfa_can_fuse_split(&pl->fwd))
      { rc = mfft_mfa_exec_split(&pl->fwd, W, pl->Y, (const limb_t *) d_i2, (uint64_t) pl->n2, p->bits1, p->j2, stream); break; }
      if (mfft_dev_split(W, pl->l, pl->pitch, (const limb_t *) d_i2, (uint64_t) pl->n2, p->bits1, p->j2, p->trunc, stream)) return MPIRFFT_ENODEV;
      rc = mfft_mfa_exec(&pl->fwd, W, pl->Y, stream);
      break;
   }
   case 2:
      if (mfft_dev_pointwise(pl->Z, pl->Y, pl->d_pw_blocks, pl->npw, pl->l, pl->pitch, stream)) return MPIRFFT_ENODEV;
      break;
   case 3:
      rc = mfft_mfa_exec(&pl->inv, pl->Z, pl->X, stream);
      break;
   case 4:
      if (mfft_dev_combine((limb_t *) d_r, (uint64_t)(pl->n1 + pl->n2), pl->X, pl->l, pl->pitch, p->bits1,
                           p->j1 + p->j2 - 1, pl->combine_work, stream)) return MPIRFFT_ENODEV;
      break;
   default: return MPIRFFT_EINVAL;
   }
   return rc;
}

/* the sharded plan on one rank: the three exchanges are local copies, the carry hand-off has nobody to
   hand to (what leaves limb n1+n2-1 is zero) */
static int exec_big(mpirfft_mul_plan *pl, mp_limb_t *d_r, const mp_limb_t *d_i1, const mp_limb_t *d_i2, void *stream)
{
   mpirfft_smul_layout lay; int which, rc; size_t tr;
   if ((rc = mpirfft_smul_info(pl->big, &lay)) != 0) return rc;
   tr = (size_t) lay.trunc_rows * lay.ncl * lay.block_limbs * sizeof(limb_t);
   for (which = 0; which < 2; which++)
   {
      if ((rc = mpirfft_smul_phase(pl->big, 0, which, which ? d_i2 : d_i1, stream)) != 0) return rc;
      if (mfft_dev_d2d(lay.recv, lay.send, tr, stream)) return MPIRFFT_ENODEV;
      if ((rc = mpirfft_smul_phase(pl->big, 1, which, NULL, stream)) != 0) return rc;
   }
   if ((rc = mpirfft_smul_phase(pl->big, 2, 0, NULL, stream)) != 0) return rc;
   if ((rc = mpirfft_smul_phase(pl->big, 3, 0, NULL, stream)) != 0) return rc;
   if (mfft_dev_d2d(lay.work, lay.send, tr, stream)) return MPIRFFT_ENODEV;
   if ((rc = mpirfft_smul_phase(pl->big, 4, 0, NULL, stream)) != 0) return rc;
   if (mfft_dev_d2d(lay.recv, lay.send, tr, stream)) return MPIRFFT_ENODEV;
   if ((rc = mpirfft_smul_phase(pl->big, 5, 0, NULL, stream)) != 0) return rc;
   if ((rc = mpirfft_smul_phase(pl->big, 6, 0, NULL, stream)) != 0) return rc;
   if (mfft_dev_d2d(d_r, lay.out, (size_t)(lay.limb_hi - lay.limb_lo) * sizeof(limb_t), stream)) return MPIRFFT_ENODEV;
   return 0;
}

int mpirfft_mul_exec_device(mpirfft_mul_plan *pl, mp_limb_t *d_r, const mp_limb_t *d_i1,
                            const mp_limb_t *d_i2, void *stream)
{
   int ph, rc;
   mfft_lock();
   if ((rc = mfft_try_device()) != 0) goto done;
   if (pl->big) { rc = exec_big(pl, d_r, d_i1, d_i2, stream); goto done; }
   if (d_i1 == d_i2 && pl->n1 == pl->n2)
   {  /* squaring: one forward transform, the spectrum multiplied by itself */
      if ((rc = exec_phase(pl, 0, d_r, d_i1, d_i2, stream)) != 0) goto done;
      if (mfft_dev_pointwise(pl->Z, pl->Z, pl->d_pw_blocks, pl->npw, pl->l, pl->pitch, stream)) { rc = MPIRFFT_ENODEV; goto done; }
      for (ph = 3; ph < 5; ph++)
         if ((rc = exec_phase(pl, ph, d_r, d_i1, d_i2, stream)) != 0) goto done;
      goto done;
   }
   if (pl->X2 && !mfft_dev_profile_is_on())     /* (per-launch event timing wants the launches one after the other) */
   {  /* fork: the second operand's transform on its own stream and slab; join before the products */
      rc = MPIRFFT_ENODEV;
      if (mfft_dev_event_record(pl->ev_fork, stream) || mfft_dev_stream_wait(pl->s_fwd2, pl->ev_fork)) goto done;
      if ((rc = exec_phase(pl, 0, d_r, d_i1, d_i2, stream)) != 0) goto done;
      if ((rc = exec_phase(pl, 1, d_r, d_i1, d_i2, pl->s_fwd2)) != 0) goto done;
      rc = MPIRFFT_ENODEV;
      if (mfft_dev_event_record(pl->ev_join, pl->s_fwd2) || mfft_dev_stream_wait(stream, pl->ev_join)) goto done;
      for (ph = 2; ph < 5; ph++)
         if ((rc = exec_phase(pl, ph, d_r, d_i1, d_i2, stream)) != 0) break;
      goto done;
   }
   for (ph = 0; ph < 5; ph++)
      if ((rc = exec_phase(pl, ph, d_r, d_i1, d_i2, stream)) != 0) break;
done:
   mfft_unlock();
   return rc;
}

/* ---- operands / result in pageable host memory: worker threads move chunks between the caller's buffers
   and the plan's pinned staging buffer, the chunks travel by DMA as they become ready (hostcopy.c) ---- */
static size_t stage_chunk(void)
{  /* 1 MB; MPIRFFT_COPY_CHUNK (bytes, a multiple of 4096) lets the tests drive many chunks through small operands */
   static size_t c = 0;
   if (!c) { const char *e = getenv("MPIRFFT_COPY_CHUNK"); long v = e ? atol(e) : 0; c = (v >= 4096 && v % 4096 == 0) ? (size_t) v : (size_t) 1 << 20; }
   return c;
}
#define STAGE_CHUNK stage_chunk()
typedef struct {
   unsigned char *stage; const unsigned char *src[2]; size_t bytes[2], nch[2];
   unsigned char *ready;                      /* per chunk: copied into the staging buffer */
} stage_in_job;

static void stage_in_chunk(void *arg, size_t i)
{
   stage_in_job *j = (stage_in_job *) arg;
   const int w = i >= j->nch[0];
   const size_t c = w ? i - j->nch[0] : i, off = c*STAGE_CHUNK;
   const size_t len = j->bytes[w] - off < STAGE_CHUNK ? j->bytes[w] - off : STAGE_CHUNK;
   memcpy(j->stage + (w ? j->bytes[0] : 0) + off, j->src[w] + off, len);
   __atomic_store_n(&j->ready[i], 1, __ATOMIC_RELEASE);
}

typedef struct { unsigned char *dst; const unsigned char *stage; size_t total, chunk; void **ev; int failed; } stage_out_job;

static void stage_out_chunk(void *arg, size_t i)
{
   stage_out_job *j = (stage_out_job *) arg;
   const size_t off = i*j->chunk, len = j->total - off < j->chunk ? j->total - off : j->chunk;
   if (mfft_dev_event_sync(j->ev[i]) != 0) { __atomic_store_n(&j->failed, 1, __ATOMIC_RELAXED); return; }
   memcpy(j->dst + off, j->stage + off, len);
}

static int stage_ready(mpirfft_mul_plan *pl, size_t bytes)
{
   int k;
   static long smin = -1;      /* below 4 MB (operands + result) the plain copies win: waking the workers costs more
                                  than the driver's own staging (measured at 2^16 x 2^16 limbs: 0.31 vs 0.43-0.56 ms) */
   if (smin < 0) { const char *e = getenv("MPIRFFT_COPY_MIN"); smin = e ? atol(e) : 4l << 20; if (smin < 0) smin = 0; }
   if (bytes > MUL_STAGE_MAX || bytes < (size_t) smin || mfft_hc_threads() <= 0) return 0;
   if (!pl->h_stage)
   {
      pl->h_stage = (unsigned char *) mfft_host_alloc_pinned(bytes);
      if (!pl->h_stage) return 0;
      pl->h_stage_bytes = bytes;
      for (k = 0; k < MUL_STAGE_EVENTS; k++)
         if (!(pl->ev_chunk[k] = mfft_dev_event_create())) return 0;
   }
   return pl->h_stage_bytes >= bytes && pl->ev_chunk[MUL_STAGE_EVENTS - 1] != NULL;
}

/* the result: D2H into the staging buffer chunk by chunk, each chunk copied out behind its event */
static int staged_result(mpirfft_mul_plan *pl, mp_limb_t *r, size_t total, void *stream)
{
   stage_out_job jo; size_t n, k;
   jo.chunk = (total + MUL_STAGE_EVENTS - 1)/MUL_STAGE_EVENTS;
   if (jo.chunk < STAGE_CHUNK) jo.chunk = STAGE_CHUNK;
   jo.chunk = (jo.chunk + 4095) & ~(size_t) 4095;
   n = (total + jo.chunk - 1)/jo.chunk;
   jo.dst = (unsigned char *) r; jo.stage = pl->h_stage; jo.total = total; jo.ev = pl->ev_chunk; jo.failed = 0;
   for (k = 0; k < n; k++)
   {
      const size_t off = k*jo.chunk, len = total - off < jo.chunk ? total - off : jo.chunk;
      if (mfft_dev_d2h(pl->h_stage + off, (const unsigned char *) pl->d_r + off, len, stream) ||
          mfft_dev_event_record(pl->ev_chunk[k], stream)) return MPIRFFT_ENODEV;
   }
   mfft_hc_begin(stage_out_chunk, &jo, n);
   mfft_hc_help();
   mfft_hc_end(n);
   if (jo.failed || mfft_dev_sync(stream)) return MPIRFFT_ENODEV;
   return 0;
}

int mpirfft_mul_exec_host(mpirfft_mul_plan *pl, mp_limb_t *r, const mp_limb_t *i1, const mp_limb_t *i2)
{
   int rc;
   size_t b1 = (size_t) pl->n1*8, b2 = (size_t) pl->n2*8;
   mfft_lock();
   if ((rc = mfft_try_device()) != 0) { mfft_unlock(); return rc; }
   if (!pl->d_i1)
   {
      pl->d_i1 = (limb_t *) mfft_dev_alloc(b1); pl->d_i2 = (limb_t *) mfft_dev_alloc(b2);
      pl->d_r = (limb_t *) mfft_dev_alloc(b1 + b2);
      if (!pl->d_i1 || !pl->d_i2 || !pl->d_r) { mfft_unlock(); return MPIRFFT_ENOMEM; }
      pl->s_copy = mfft_dev_stream_create(); pl->s_comp = mfft_dev_stream_create();
      pl->ev1 = mfft_dev_event_create(); pl->ev2 = mfft_dev_event_create();
      if (!pl->s_copy || !pl->s_comp || !pl->ev1 || !pl->ev2) { mfft_unlock(); return MPIRFFT_ENODEV; }
   }
   /* the second operand travels while the first one is being split and transformed */
   rc = MPIRFFT_ENODEV;
   {
      /* Pageable operands (what a caller that merely swaps libraries hands in): a plain copy blocks the host
         while the driver stages it at 10-13 GB/s.  Worker threads copy 1 MB chunks into the pinned staging
         buffer instead and this thread issues each chunk's DMA as soon as it is ready; the second operand's
         chunks are prepared while the first transform is already running.  Pinned operands go as they are. */
      const int in_staged = !(mfft_dev_host_is_pinned(i1) && mfft_dev_host_is_pinned(i2)) && stage_ready(pl, b1 + b2);
      const int out_staged = !mfft_dev_host_is_pinned(r) && stage_ready(pl, b1 + b2);
      stage_in_job ji; size_t nin = 0, k;
      memset(&ji, 0, sizeof ji);
      if (in_staged)
      {
         ji.stage = pl->h_stage; ji.src[0] = (const unsigned char *) i1; ji.src[1] = (const unsigned char *) i2;
         ji.bytes[0] = b1; ji.bytes[1] = b2;
         ji.nch[0] = (b1 + STAGE_CHUNK - 1)/STAGE_CHUNK; ji.nch[1] = (b2 + STAGE_CHUNK - 1)/STAGE_CHUNK;
         nin = ji.nch[0] + ji.nch[1];
         ji.ready = (unsigned char *) calloc(nin ? nin : 1, 1);
         if (!ji.ready) { rc = MPIRFFT_ENOMEM; goto done; }
         mfft_hc_begin(stage_in_chunk, &ji, nin);
      }
#define STAGED_H2D(w, dptr)                                                                                   \
      for (k = 0; k < ji.nch[w]; k++)                                                                          \
      {                                                                                                        \
         const size_t off = k*STAGE_CHUNK, len = ji.bytes[w] - off < STAGE_CHUNK ? ji.bytes[w] - off : STAGE_CHUNK; \
         while (!__atomic_load_n(&ji.ready[(w ? ji.nch[0] : 0) + k], __ATOMIC_ACQUIRE)) sched_yield();          \
         if (mfft_dev_h2d((unsigned char *)(dptr) + off, pl->h_stage + (w ? b1 : 0) + off, len, pl->s_copy)) goto fail_in; \
      }
      if (pl->big)
      {  /* rings above 512 limbs: the sharded plan on one rank runs as a whole behind both copies */
         if (in_staged) { STAGED_H2D(0, pl->d_i1) STAGED_H2D(1, pl->d_i2) mfft_hc_end(nin); free(ji.ready); ji.ready = NULL; }
         else if (mfft_dev_h2d(pl->d_i1, i1, b1, pl->s_copy) || mfft_dev_h2d(pl->d_i2, i2, b2, pl->s_copy)) goto fail_in;
         if (mfft_dev_event_record(pl->ev1, pl->s_copy) || mfft_dev_stream_wait(pl->s_comp, pl->ev1)) goto fail_in;
         if ((rc = exec_big(pl, (mp_limb_t *) pl->d_r, (const mp_limb_t *) pl->d_i1, (const mp_limb_t *) pl->d_i2, pl->s_comp)) != 0) goto done;
         if (out_staged) { rc = staged_result(pl, r, b1 + b2, pl->s_comp); goto done; }
         rc = MPIRFFT_ENODEV;
         if (mfft_dev_d2h(r, pl->d_r, b1 + b2, pl->s_comp) || mfft_dev_sync(pl->s_comp)) goto done;
         rc = 0;
         goto done;
      }
      /* The first transform is launched BEFORE the second copy is issued: the GPU should be busy with operand 1
         while operand 2 is on its way */
      if (in_staged) { STAGED_H2D(0, pl->d_i1) }
      else if (mfft_dev_h2d(pl->d_i1, i1, b1, pl->s_copy)) goto fail_in;
      if (mfft_dev_event_record(pl->ev1, pl->s_copy) || mfft_dev_stream_wait(pl->s_comp, pl->ev1)) goto fail_in;
      if ((rc = mpirfft_mul_exec_phase(pl, 0, (mp_limb_t *) pl->d_r, (const mp_limb_t *) pl->d_i1, (const mp_limb_t *) pl->d_i2, pl->s_comp)) != 0) goto fail_in;
      rc = MPIRFFT_ENODEV;
      if (in_staged) { STAGED_H2D(1, pl->d_i2) }
      else if (mfft_dev_h2d(pl->d_i2, i2, b2, pl->s_copy)) goto fail_in;
#undef STAGED_H2D
      if (mfft_dev_event_record(pl->ev2, pl->s_copy)) goto fail_in;
      if (in_staged) { mfft_hc_end(nin); free(ji.ready); ji.ready = NULL; }
      if (pl->X2)
      {  /* the second transform starts on its own stream as soon as its operand has arrived */
         if (mfft_dev_stream_wait(pl->s_fwd2, pl->ev2)) goto done;
         if ((rc = mpirfft_mul_exec_phase(pl, 1, (mp_limb_t *) pl->d_r, (const mp_limb_t *) pl->d_i1, (const mp_limb_t *) pl->d_i2, pl->s_fwd2)) != 0) goto done;
         rc = MPIRFFT_ENODEV;
         if (mfft_dev_event_record(pl->ev_join, pl->s_fwd2) || mfft_dev_stream_wait(pl->s_comp, pl->ev_join)) goto done;
      }
      else if (mfft_dev_stream_wait(pl->s_comp, pl->ev2)) goto done;
      {
         int ph;
         for (ph = pl->X2 ? 2 : 1; ph < 5; ph++)
            if ((rc = mpirfft_mul_exec_phase(pl, ph, (mp_limb_t *) pl->d_r, (const mp_limb_t *) pl->d_i1, (const mp_limb_t *) pl->d_i2, pl->s_comp)) != 0) goto done;
      }
      if (out_staged) { rc = staged_result(pl, r, b1 + b2, pl->s_comp); goto done; }
      rc = MPIRFFT_ENODEV;
      if (mfft_dev_d2h(r, pl->d_r, b1 + b2, pl->s_comp) || mfft_dev_sync(pl->s_comp)) goto done;
      rc = 0;
      goto done;
fail_in:
      /* the workers still hold pointers into ji: let them finish before it goes out of scope */
      if (ji.ready) { mfft_hc_end(nin); free(ji.ready); }
      mfft_dev_sync(pl->s_copy);
   }
done:
   mfft_unlock();
   return rc;
}

/* a tiny plan cache so that the drop-in symbol does not rebuild schedules on every call */
#define PLAN_CACHE 4
static mpirfft_mul_plan *g_cache[PLAN_CACHE];
static unsigned g_cache_next = 0;

static void mul_dropin(const char *fn, int sqrt2, mp_limb_t *r1, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2,
                       mp_bitcnt_t depth, mp_bitcnt_t w);

void new_mpn_mul(mp_limb_t *r1, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2,
                 mp_bitcnt_t depth, mp_bitcnt_t w)
{ mul_dropin("new_mpn_mul", 0, r1, i1, n1, i2, n2, depth, w); }

/* mul_fft.c:3573  the same product through the sqrt2 transforms of length 4n */
void new_mpn_mul6(mp_limb_t *r1, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2,
                  mp_bitcnt_t depth, mp_bitcnt_t w)
{ mul_dropin("new_mpn_mul6", 1, r1, i1, n1, i2, n2, depth, w); }

static void mul_dropin(const char *fn, int sqrt2, mp_limb_t *r1, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2,
                       mp_bitcnt_t depth, mp_bitcnt_t w)
{
   mpirfft_mul_plan *pl = NULL; int k, rc;
   mpirfft_mul_params p;
   if (!r1 || !i1 || !i2) mfft_die(fn, "null operand");
   if ((sqrt2 ? mpirfft_mul6_params_get(&p, n1, n2, depth, w) : mpirfft_mul_params_get(&p, n1, n2, depth, w)) != 0)
      mfft_die(fn, "illegal parameters n1=%ld n2=%ld depth=%lu w=%lu: need 64 | 2^depth*w and %s "
               "(mul_fft.c:3186-3187)", (long) n1, (long) n2, (unsigned long) depth, (unsigned long) w,
               sqrt2 ? "2^(depth+1) < j1+j2-1 <= 2^(depth+2)" : "j1+j2-1 <= 2^(depth+1)");
   /* the lock is held across lookup, eviction and execution: a concurrent call with another shape
      cannot destroy the plan this thread is executing (the reference is re-entrant) */
   mfft_lock(); mfft_require_device(fn);
   for (k = 0; k < PLAN_CACHE; k++)
      if (g_cache[k] && g_cache[k]->n1 == n1 && g_cache[k]->n2 == n2 && g_cache[k]->depth == depth && g_cache[k]->w == w &&
          g_cache[k]->sqrt2 == sqrt2)
         pl = g_cache[k];
   if (!pl)
   {
      if ((rc = plan_create(&pl, n1, n2, depth, w, sqrt2)) != 0)
         mfft_die(fn, "cannot build the plan (code %d): %s", rc, mfft_dev_last_error());
      k = (int)(g_cache_next++ % PLAN_CACHE);
      if (g_cache[k]) mpirfft_mul_plan_destroy(g_cache[k]);
      g_cache[k] = pl;
   }
   if ((rc = mpirfft_mul_exec_host(pl, r1, i1, i2)) != 0)
      mfft_die(fn, "device execution failed (code %d): %s", rc, mfft_dev_last_error());
   mfft_unlock();
}

/* mpn_mul-shaped entry (what the FIXME at mul_fft.c:3177-3178 asks for): chooses (depth, w) itself */
void mpirfft_mpn_mul(mp_limb_t *r, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2)
{
   mp_bitcnt_t depth, w; int sqrt2;
   if (n1 <= 0 || n2 <= 0) mfft_die("mpirfft_mpn_mul", "operand sizes must be positive");
   if (mpirfft_choose_params6(n1, n2, &depth, &w, &sqrt2) != 0)
      mfft_die("mpirfft_mpn_mul", "no legal (depth, w) for %ld x %ld limbs", (long) n1, (long) n2);
   if (sqrt2) new_mpn_mul6(r, i1, n1, i2, n2, depth, w);
   else new_mpn_mul(r, i1, n1, i2, n2, depth, w);
}
