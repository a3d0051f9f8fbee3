/* xform.c -- see xform.h */
#define _POSIX_C_SOURCE 200809L
#include "xform.h"
#include "../../../include/mpirfft_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NW_of(s) ((s)->NW)

int mfft_xform_halves(const mfft_xform *x) { return x->fused ? 1 : 2; }

/* Shape of a slot-view op for the carry-save stage kernel (mfft_cs_stage.h): 0 if the kernel can
 * run it (kind / kparam filled in), -1 if not (then the whole transform keeps the ballot-carry
 * stage kernel). */
static int fold128(int sign, uint32_t e, uint64_t NW, uint32_t *yc, uint32_t *neg)
{
   if (e >= NW) { e -= (uint32_t) NW; sign = -sign; }
   if (e % 128) return -1;
   *yc = e/128; *neg = (sign < 0);
   return 0;
}

static int classify_slot_op(mfft_op *o, uint64_t NW)
{
   const int hasB = (o->inB != MFFT_NONE), hasT = (o->outT != MFFT_NONE);
   uint32_t ySA = 0, nSA = 0, ySB = 0, nSB = 0, yTA = 0, nTA = 0, yTB = 0, nTB = 0;
   int al;
   o->kind = MFFT_K_ANY; o->kparam = 0;
   if (!o->sSA) return -1;
   if (o->cSA | o->cSB | o->cTA | o->cTB) return (!hasB && !hasT) ? 0 : -1;      /* twisted: one operand only */
   if (!hasB && !hasT && o->sSA == 1 && o->eSA == 1) { o->kind = MFFT_K_DBL; return 0; }
   if (!hasB && !hasT && o->sSA == 1 && o->eSA > 2*NW - 32 && o->eSA < 2*NW) { o->kind = MFFT_K_SHR; o->kparam = (uint32_t)(2*NW - o->eSA); return 0; }
   if (hasB && !hasT && o->sSA == 1 && o->sSB == 1 && o->eSA == 2*NW - 1 && o->eSB == 2*NW - 1) { o->kind = MFFT_K_HALF; return 0; }
   if (hasB && o->sSA == 1 && o->eSA == 1 && o->sSB == -1 && o->eSB == 0)
   {  /* S = 2A - B, optionally T = +-(A - B) rotated */
      if (!hasT) { o->kind = MFFT_K_2AMB; return 0; }
      if (!o->sTA || !o->sTB || fold128(o->sTA, o->eTA, NW, &yTA, &nTA) || fold128(o->sTB, o->eTB, NW, &yTB, &nTB)) return -1;
      if (yTA != yTB || nTA == nTB) return -1;
      o->kind = MFFT_K_2AMB; o->kparam = yTA | (nTA << 31);
      return 0;
   }
   al = fold128(o->sSA, o->eSA, NW, &ySA, &nSA) == 0;
   if (al && hasB) al = o->sSB && fold128(o->sSB, o->eSB, NW, &ySB, &nSB) == 0;
   if (al && hasT) al = o->sTA && fold128(o->sTA, o->eTA, NW, &yTA, &nTA) == 0;
   if (al && hasT && hasB) al = o->sTB && fold128(o->sTB, o->eTB, NW, &yTB, &nTB) == 0;
   if (!al) return (!hasB && !hasT) ? 0 : -1;                                      /* any rotation of one operand */
   if (hasB && hasT)
   {
      if (ySA == 0 && !nSA && ySB == 0 && !nSB && yTA == yTB && nTA != nTB) { o->kind = MFFT_K_FWD; o->kparam = yTA | (nTA << 31); return 0; }
      if (ySA == 0 && !nSA && yTA == 0 && !nTA && ySB == yTB && nSB != nTB) { o->kind = MFFT_K_INV; o->kparam = ySB | (nSB << 31); return 0; }
      return -1;
   }
   if (!hasB && !hasT) { o->kind = MFFT_K_ROT; o->kparam = ySA | (nSA << 31); return 0; }
   if (hasB && !hasT && ySA == 0 && !nSA && ySB == 0 && !nSB) { o->kind = MFFT_K_ADD; return 0; }
   return -1;
}

void mfft_xform_free(mfft_xform *x)
{
   uint32_t i;
   if (x->dp)
   {
      for (i = 0; i < x->P.npasses; i++)
      { mfft_dev_free(x->dp[i].d_tiles); mfft_dev_free(x->dp[i].d_pos); mfft_dev_free(x->dp[i].d_ops); mfft_dev_free(x->dp[i].d_stoff); }
      free(x->dp);
   }
   mfft_passes_free(&x->P);
   if (x->ds.s) { mfft_dsched_free(&x->ds); x->s = NULL; }
   if (x->s) mfft_sched_free(x->s);
   free(x->h_batch);
   mfft_dev_free(x->d_cw);
   mfft_dev_free(x->d_batch); mfft_dev_free(x->d_dst_base); mfft_dev_free(x->d_dstpos); mfft_dev_free(x->d_moves);
   memset(x, 0, sizeof(*x));
}

int mfft_xform_build(mfft_xform *x, mfft_sched *s, uint32_t l, uint32_t slot_stride, uint64_t half_blocks,
                     const mfft_batch *batch, uint32_t nbatch, const uint32_t *dst_of, uint32_t nout,
                     const uint32_t *dst_base, uint32_t dst_stride, uint32_t shift, int normalise)
{
   const char *env = getenv("MPIRFFT_UNFUSED");
   uint32_t S = s->S, k; int rc = MPIRFFT_ENOMEM;
   memset(x, 0, sizeof(*x));
   x->s = s; x->S = S; x->nbatch = nbatch; x->nout = nout; x->dst_stride = dst_stride;
   x->shift = (uint32_t)(shift % (128ull*l)); x->normalise = normalise;
   x->g.S = S; x->g.slot_stride = slot_stride; x->g.half_blocks = half_blocks; x->g.l = l; x->g.pitch = mfft_pitch(l);
   x->fused = mfft_dev_tiles_supported(l) && !(env && env[0] == '1');
   x->d_batch = (mfft_batch *) mfft_upload(batch, sizeof(mfft_batch)*nbatch);
   x->h_batch = (mfft_batch *) malloc(sizeof(mfft_batch)*(nbatch ? nbatch : 1));
   if (x->h_batch) memcpy(x->h_batch, batch, sizeof(mfft_batch)*nbatch);
   x->d_dst_base = (uint32_t *) mfft_upload(dst_base, sizeof(uint32_t)*nbatch);
   if (!x->d_batch || !x->d_dst_base) { rc = MPIRFFT_ENODEV; goto fail; }
   if (x->fused)
   {
      uint8_t *must = (uint8_t *) calloc(S, 1);
      uint32_t *dstpos = (uint32_t *) malloc(sizeof(uint32_t)*S);
      if (!must || !dstpos) { free(must); free(dstpos); goto fail; }
      for (k = 0; k < S; k++) dstpos[k] = MFFT_NONE;
      if (x->shift)
         for (k = 0; k < nout; k++) mfft_sched_emit_op(s, k, MFFT_NONE, k, 1, x->shift, 0, 0, MFFT_NONE, 0, 0, 0, 0);
      for (k = 0; k < nout; k++) { must[s->phys[k]] = 1; dstpos[s->phys[k]] = dst_of[k]; }
      if (mfft_passes_build(&x->P, s, mfft_dev_tiles_max_npos(l), must, must) != 0) { free(must); free(dstpos); goto fail; }
      x->d_dstpos = (uint32_t *) mfft_upload(dstpos, sizeof(uint32_t)*S);
      free(must); free(dstpos);
      x->dp = (struct mfft_dpass *) calloc(x->P.npasses ? x->P.npasses : 1, sizeof(*x->dp));
      if (!x->d_dstpos || !x->dp) goto fail;
      for (k = 0; k < x->P.npasses; k++)
      {
         const mfft_pass *p = &x->P.pass[k];
         x->dp[k].d_tiles = (mfft_tile *) mfft_upload(p->tiles, sizeof(mfft_tile)*(p->ntiles ? p->ntiles : 1));
         x->dp[k].d_pos = (uint32_t *) mfft_upload(p->pos, sizeof(uint32_t)*(p->npos_total ? p->npos_total : 1));
         x->dp[k].d_ops = (mfft_tileop *) mfft_upload(p->ops, sizeof(mfft_tileop)*(p->nops_total ? p->nops_total : 1));
         x->dp[k].d_stoff = (uint32_t *) mfft_upload(p->stoff, sizeof(uint32_t)*(p->nstoff ? p->nstoff : 1));
         if (!x->dp[k].d_tiles || !x->dp[k].d_pos || !x->dp[k].d_ops || !x->dp[k].d_stoff) { rc = MPIRFFT_ENODEV; goto fail; }
      }
   } else
   {
      mfft_move *mv;
      /* carry-save stage kernel where every op has a shape it knows (even l, aligned blocks): the
         final scaling becomes one more layer of ops, the gather only resolves and normalises */
      {
         const char *e2 = getenv("MPIRFFT_NO_CS_STAGE");
         size_t i2; int ok = !(e2 && e2[0] == '1') && (l % 2 == 0) && l >= 64 && NW_of(s) % 128 == 0;
         if (ok && x->shift)
            for (k = 0; k < nout; k++) mfft_sched_emit_op(s, k, MFFT_NONE, k, 1, x->shift, 0, 0, MFFT_NONE, 0, 0, 0, 0);
         for (i2 = 0; ok && i2 < s->nops; i2++) if (classify_slot_op(&s->ops[i2], NW_of(s)) != 0) ok = 0;
         if (ok)
         {
            x->cs = 1; x->shift = 0;
            x->d_cw = (int32_t *) mfft_dev_alloc((size_t) 2*half_blocks*(l/2)*sizeof(int32_t));
            if (!x->d_cw) { rc = MPIRFFT_ENODEV; goto fail; }
         } else if (x->shift && !(e2 && e2[0] == '1') && (l % 2 == 0) && l >= 64 && NW_of(s) % 128 == 0)
         {  /* the scaling ops were emitted but the transform stays on the old path: they do the scaling there too */
            x->shift = 0;
         }
      }
      mv = (mfft_move *) malloc(sizeof(mfft_move)*(nout ? nout : 1));
      if (!mv) goto fail;
      for (k = 0; k < nout; k++) { mv[k].src_slot = s->slot[k]; mv[k].dst_pos = dst_of[k]; }
      x->d_moves = (mfft_move *) mfft_upload(mv, sizeof(mfft_move)*(nout ? nout : 1));
      free(mv);
      if (!x->d_moves) { rc = MPIRFFT_ENODEV; goto fail; }
      if (mfft_dsched_upload(&x->ds, s) != 0) { rc = MPIRFFT_ENODEV; goto fail; }
   }
   if (getenv("MPIRFFT_VERBOSE"))
      fprintf(stderr, "mpirfft xform: l=%u S=%u batch=%u -> %s\n", l, S, nbatch,
              x->fused ? "fused tiles" : (x->cs ? "carry-save stage kernel" : "ballot-carry stage kernel"));
   return 0;
fail:
   mfft_xform_free(x);
   return rc;
}

int mfft_xform_exec(const mfft_xform *x, limb_t *slab, limb_t *dst, void *stream)
{
   uint32_t i;
   if (x->fused)
   {
      for (i = 0; i < x->P.npasses; i++)
      {
         const mfft_pass *p = &x->P.pass[i];
         const int lastp = (i + 1 == x->P.npasses);
         if (mfft_dev_run_tiles(slab, &x->g, x->dp[i].d_tiles, p->ntiles, x->dp[i].d_pos, x->dp[i].d_ops, p->max_npos,
                                p->max_nops, x->d_batch, x->nbatch, lastp ? dst : NULL, x->d_dstpos, x->d_dst_base,
                                x->dst_stride, lastp ? x->normalise : 0, x->dp[i].d_stoff, (5*p->nany > p->nops_total) || p->nr4, p->tiles, p->pos, p->stoff, x->h_batch, NULL, stream) != 0) return MPIRFFT_ENODEV;
      }
      return 0;
   }
   if (x->cs)
   {
      const mfft_sched *sc = x->ds.s; uint32_t st;
      if (mfft_dev_cs_init(slab, x->d_cw, &x->g, x->d_batch, x->nbatch, stream) != 0) return MPIRFFT_ENODEV;
      for (st = 1; st <= sc->nstages; st++)
      {
         const uint32_t lo = sc->stage_off[st - 1], hi = sc->stage_off[st];
         if (hi > lo && mfft_dev_run_stage_cs(slab, x->d_cw, &x->g, x->ds.d_ops + lo, hi - lo, x->d_batch, x->nbatch, stream) != 0)
            return MPIRFFT_ENODEV;
      }
      if (mfft_dev_finalize_cs(dst, x->dst_stride, x->d_dst_base, slab, x->d_cw, &x->g, x->d_moves, x->nout, x->d_batch,
                               x->nbatch, x->normalise, stream) != 0) return MPIRFFT_ENODEV;
      return 0;
   }
   if (mfft_dsched_run(&x->ds, slab, &x->g, x->d_batch, x->nbatch, stream) != 0) return MPIRFFT_ENODEV;
   if (mfft_dev_finalize(dst, x->dst_stride, x->d_dst_base, slab, &x->g, x->d_moves, x->nout, x->d_batch, x->nbatch,
                         x->shift, x->normalise, stream) != 0) return MPIRFFT_ENODEV;
   return 0;
}
