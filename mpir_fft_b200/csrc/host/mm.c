/* mm.c -- batched multiplication mod 2^(64 L) + 1 through a negacyclic transform over a smaller
 * ring: the reference's FFT_mulmod_2expp1 (mul_fft.c:2998-3117) and the parameter heuristic of
 * fft_mulmod_2expp1 (3125-3167), for `count` independent products at once (BASELINE configs[3]:
 * 4096 products of 16 384 limbs), everything resident in HBM:
 *
 *   split      K = 2n pieces of pl = L/K limbs per operand                    3038-3040, 3049-3051
 *   forward    negacyclic transforms of length K over Z/(2^(n w) + 1)          3046, 3056
 *   pointwise  K products mod 2^(n w) + 1 per pair                             3058-3063
 *   inverse    negacyclic inverse, scaled by 2^-(depth+1), normalised          3065, 3073-3074
 *   finish     low-limb convolution mod 2^64, correction, recombination,
 *              wrap-around and final normalisation                             3067-3115
 *
 * Supported shapes: n w = 128 pl (coefficients are exactly two pieces wide, the shape the
 * reference's heuristic always produces), w even, pl a multiple of 32 limbs.
 */
#define _POSIX_C_SOURCE 200809L
#include "xform.h"
#include "../../../include/mpirfft_b200.h"
#include <stdlib.h>
#include <string.h>

struct mpirfft_mulmod_plan {
   uint32_t L, count, depth, K, l, pitch, pl; uint64_t w;
   mfft_xform fwd, inv;
   limb_t *W, *Ah, *Bh, *Cc;
   uint32_t *d_idx;
   size_t dev_bytes;
};

int mpirfft_mulmod_params(mp_size_t r_limbs, mp_bitcnt_t *depth, mp_bitcnt_t *w)
{
   /* inner ring of 128 limbs when the size allows, else 64, 256, 512: K = 2L/l pieces, n = K/2,
      w = 64 l / n.  Measured for 4096 x 16 384 limbs on B200: 19.2 / 18.9 / 23.2 / 41.1 ms per batch
      with l = 128 / 64 / 256 / 512 -- the pointwise work falls with l, the transform work grows */
   static const uint32_t cand[4] = { 128, 64, 256, 512 };
   const char *env = getenv("MPIRFFT_MM_INNER");      /* developer aid: force the inner ring (limbs) */
   int i;
   if (r_limbs < 64) return MPIRFFT_EINVAL;
   for (i = 0; i < 4; i++)
   {
      uint64_t l = cand[i], K, n, ww; uint32_t d = 0;
      if (env && env[0] && (uint64_t) atoi(env) != l) continue;
      if ((2*(uint64_t) r_limbs) % l) continue;
      K = 2*(uint64_t) r_limbs/l;
      if (K < 4 || (K & (K - 1))) continue;
      n = K/2; if ((64*l) % n) continue;
      ww = 64*l/n; if (ww & 1) continue;
      if ((l/2) % 32) continue;
      while (((uint64_t)1 << d) < n) d++;
      *depth = d; *w = ww; return 0;
   }
   return MPIRFFT_EINVAL;
}

void mpirfft_mulmod_plan_destroy(mpirfft_mulmod_plan *pl)
{
   if (!pl) return;
   mfft_lock();
   mfft_xform_free(&pl->fwd); mfft_xform_free(&pl->inv);
   mfft_dev_free(pl->W); mfft_dev_free(pl->Ah); mfft_dev_free(pl->Bh); mfft_dev_free(pl->Cc); mfft_dev_free(pl->d_idx);
   mfft_unlock();
   free(pl);
}

int mpirfft_mulmod_plan_create(mpirfft_mulmod_plan **out, mp_size_t r_limbs, mp_bitcnt_t depth, mp_bitcnt_t w, size_t count)
{
   mpirfft_mulmod_plan *pl; uint64_t n, NW, K, i; int rc = MPIRFFT_EINVAL;
   mfft_sched *s = NULL; mfft_batch *b = NULL; uint32_t *dst_of = NULL, *dst_base = NULL, *idx = NULL;
   size_t nblk, bb;
   *out = NULL;
   if (r_limbs <= 0 || depth < 1 || depth > 12 || w == 0 || (w & 1) || count == 0 || count > 0x100000) return MPIRFFT_EINVAL;
   n = (uint64_t)1 << depth; K = 2*n; NW = n*w;
   if (NW % 64 || ((uint64_t) r_limbs) % K) return MPIRFFT_EINVAL;
   if (NW != 128*((uint64_t) r_limbs/K) || ((uint64_t) r_limbs/K) % 32) return MPIRFFT_EINVAL;
   if ((uint64_t) count*K > 0x7fffffffu) return MPIRFFT_EINVAL;
   pl = (mpirfft_mulmod_plan *) calloc(1, sizeof(*pl));
   if (!pl) return MPIRFFT_ENOMEM;
   pl->L = (uint32_t) r_limbs; pl->count = (uint32_t) count; pl->depth = (uint32_t) depth; pl->w = w;
   pl->K = (uint32_t) K; pl->l = (uint32_t)(NW/64); pl->pitch = mfft_pitch(pl->l); pl->pl = (uint32_t)(r_limbs/K);
   mfft_lock();
   if ((rc = mfft_try_device()) != 0) goto fail;
   rc = MPIRFFT_ENOMEM;
   nblk = (size_t) count*K;
   b = (mfft_batch *) calloc(count, sizeof(mfft_batch));
   dst_of = (uint32_t *) malloc(sizeof(uint32_t)*K);
   dst_base = (uint32_t *) malloc(sizeof(uint32_t)*count);
   idx = (uint32_t *) malloc(sizeof(uint32_t)*nblk);
   if (!b || !dst_of || !dst_base || !idx) goto fail;
   for (i = 0; i < count; i++) { b[i].base = (uint32_t)(i*K); b[i].parity = 0; b[i].col = 0; dst_base[i] = (uint32_t)(i*K); }
   for (i = 0; i < K; i++) dst_of[i] = (uint32_t) i;
   for (i = 0; i < nblk; i++) idx[i] = (uint32_t) i;
   /* forward: outputs normalised for the pointwise products (3047, 3060) */
   if (!(s = mfft_sched_new((uint32_t) K, NW))) goto fail;
   if (mfft_sched_emit(s, MFFT_T_FFT_NEGACYCLIC, 0, 1, n, w, 0, 0, 0, 0) != 0) { rc = MPIRFFT_EINVAL; goto fail; }
   rc = mfft_xform_build(&pl->fwd, s, pl->l, 1, nblk, b, (uint32_t) count, dst_of, (uint32_t) K, dst_base, 1, 0, 1);
   s = NULL;
   if (rc != 0) goto fail;
   /* inverse: scaled by 2^-(depth+1) and normalised (3073-3074) */
   rc = MPIRFFT_ENOMEM;
   if (!(s = mfft_sched_new((uint32_t) K, NW))) goto fail;
   if (mfft_sched_emit(s, MFFT_T_IFFT_NEGACYCLIC, 0, 1, n, w, 0, 0, 0, 0) != 0) { rc = MPIRFFT_EINVAL; goto fail; }
   rc = mfft_xform_build(&pl->inv, s, pl->l, 1, nblk, b, (uint32_t) count, dst_of, (uint32_t) K, dst_base, 1,
                         (uint32_t)(2*NW - (depth + 1)), 1);
   s = NULL;
   if (rc != 0) goto fail;
   rc = MPIRFFT_ENOMEM;
   bb = (size_t) pl->pitch*sizeof(limb_t);
   pl->W  = (limb_t *) mfft_dev_alloc((size_t) mfft_xform_halves(&pl->fwd)*nblk*bb);
   pl->Ah = (limb_t *) mfft_dev_alloc((size_t) mfft_xform_halves(&pl->inv)*nblk*bb);
   pl->Bh = (limb_t *) mfft_dev_alloc(nblk*bb);
   pl->Cc = (limb_t *) mfft_dev_alloc(nblk*bb);
   pl->d_idx = (uint32_t *) mfft_upload(idx, sizeof(uint32_t)*nblk);
   if (!pl->W || !pl->Ah || !pl->Bh || !pl->Cc || !pl->d_idx) goto fail;
   pl->dev_bytes = ((size_t) mfft_xform_halves(&pl->fwd) + mfft_xform_halves(&pl->inv) + 2)*nblk*bb;
   free(b); free(dst_of); free(dst_base); free(idx);
   mfft_unlock();
   *out = pl;
   return 0;
fail:
   if (s) mfft_sched_free(s);
   free(b); free(dst_of); free(dst_base); free(idx);
   mfft_unlock();
   mpirfft_mulmod_plan_destroy(pl);
   return rc;
}

size_t mpirfft_mulmod_plan_device_bytes(const mpirfft_mulmod_plan *pl) { return pl->dev_bytes; }

/* r[k] = a[k]*b[k] mod 2^(64 L)+1 for k < count (<= the plan's count): blocks of L+1 limbs at the
 * given pitch, canonical; r may alias a.  phase < 0: everything; 0..4: one step (bench.py). */
int mpirfft_mulmod_plan_exec(mpirfft_mulmod_plan *pl, mp_limb_t *d_r, const mp_limb_t *d_a, const mp_limb_t *d_b,
                             size_t pitch, int phase, void *stream)
{
   uint32_t n = pl->count;
   if (pitch < (size_t) pl->L + 1 || pitch > 0xffffffffu) return MPIRFFT_EINVAL;
   if (phase < 0 || phase == 0)
   {
      if (mfft_dev_mm_split(pl->W, pl->l, pl->pitch, (const limb_t *) d_a, (uint32_t) pitch, pl->K, pl->pl, n, stream)) return MPIRFFT_ENODEV;
      if (mfft_xform_exec(&pl->fwd, pl->W, pl->Ah, stream)) return MPIRFFT_ENODEV;
   }
   if (phase < 0 || phase == 1)
   {
      if (mfft_dev_mm_split(pl->W, pl->l, pl->pitch, (const limb_t *) d_b, (uint32_t) pitch, pl->K, pl->pl, n, stream)) return MPIRFFT_ENODEV;
      if (mfft_xform_exec(&pl->fwd, pl->W, pl->Bh, stream)) return MPIRFFT_ENODEV;
   }
   if (phase < 0 || phase == 2)
      if (mfft_dev_pointwise(pl->Ah, pl->Bh, pl->d_idx, n*pl->K, pl->l, pl->pitch, stream)) return MPIRFFT_ENODEV;
   if (phase < 0 || phase == 3)
      if (mfft_xform_exec(&pl->inv, pl->Ah, pl->Cc, stream)) return MPIRFFT_ENODEV;
   if (phase < 0 || phase == 4)
      if (mfft_dev_mm_finish((limb_t *) d_r, (uint32_t) pitch, (const limb_t *) d_a, (const limb_t *) d_b, (uint32_t) pitch,
                             pl->Cc, pl->l, pl->pitch, pl->K, pl->pl, n, stream)) return MPIRFFT_ENODEV;
   return 0;
}
