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


/* ---- big rings: in place, several layers per pass on chunk slices --------------------------------
 * Position view of the schedule (ops on physical positions of slab half 0, stages under in-place
 * hazards).  A stage whose ops all rotate by whole chunks is slice-local: with G = gcd of its
 * rotations (in chunks) the chunks { i0 + G t } of all coefficients form closed sub-problems.
 * Consecutive slice-local stages are merged into one pass as long as the positions they connect
 * fit a shared-memory tile of slices; the other stages (bit-granular twist rotations, halving,
 * final scaling) run one per launch with their operands staged whole (k_stage_cs_ip). */
static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

static int cmp_pstage_op(const void *x, const void *y)
{
   const mfft_op *a = (const mfft_op *) x, *b = (const mfft_op *) y;
   if (a->pstage != b->pstage) return a->pstage < b->pstage ? -1 : 1;
   return 0;
}

static int op_slice_local(const mfft_op *o)
{ return o->kind == MFFT_K_FWD || o->kind == MFFT_K_INV || o->kind == MFFT_K_ROT || o->kind == MFFT_K_ADD ||
         o->kind == MFFT_K_2AMB || o->kind == MFFT_K_DBL; }

/* exponents that are whole chunks are divided by the slice stride: the slice is a ring of its own */
static uint32_t scale_e(uint32_t e, uint32_t gs) { return (e % 128u) ? e : (e / 128u / gs) * 128u; }

static int big_build(mfft_xform *x, const mfft_sched *s, uint32_t l, const uint8_t *live_out)
{
   const uint32_t S = s->S, NCH = l/2, maxst = s->npstages;
   static const uint32_t nchv_ok[6] = { 32, 64, 96, 128, 192, 256 };
   mfft_op *ops = NULL, *win = NULL; size_t *st_off = NULL; uint32_t *last_read = NULL, *stage_g = NULL;
   size_t k; uint32_t t, cur; int rc = -1;
   const size_t budget = 96 * 1024;                /* two CTAs per SM */
   if (NCH % 32 || !maxst) return -1;
   ops = (mfft_op *) malloc(sizeof(mfft_op) * (s->nops ? s->nops : 1));
   win = (mfft_op *) malloc(sizeof(mfft_op) * (s->nops ? s->nops : 1));
   st_off = (size_t *) calloc((size_t) maxst + 2, sizeof(size_t));
   last_read = (uint32_t *) calloc(S ? S : 1, sizeof(uint32_t));
   stage_g = (uint32_t *) calloc((size_t) maxst + 2, sizeof(uint32_t));      /* 0: not slice-local */
   x->bp = (mfft_bigpass *) calloc((size_t) maxst + 1, sizeof(mfft_bigpass));
   if (!ops || !win || !st_off || !last_read || !stage_g || !x->bp) goto done;
   memcpy(ops, s->ops, sizeof(mfft_op) * s->nops);
   qsort(ops, s->nops, sizeof(mfft_op), cmp_pstage_op);
   for (k = 0; k < s->nops; k++)
   {
      const mfft_op *o = &ops[k];
      st_off[o->pstage]++;
      if (o->pstage > last_read[o->pA]) last_read[o->pA] = o->pstage;
      if (o->pB != MFFT_NONE && o->pstage > last_read[o->pB]) last_read[o->pB] = o->pstage;
   }
   { size_t acc = 0; for (t = 0; t <= maxst + 1; t++) { size_t c = st_off[t]; st_off[t] = acc; acc += c; } }
   for (t = 1; t <= maxst; t++)
   {
      uint64_t G = NCH; int ok = 1;
      for (k = st_off[t]; ok && k < st_off[t + 1]; k++)
      {
         if (!op_slice_local(&ops[k])) ok = 0;
         else G = gcd64(G, ops[k].kparam & 0x7fffffffu);      /* gcd(G, 0) = G */
      }
      stage_g[t] = ok ? (uint32_t) G : 0;
   }
   cur = 1;
   while (cur <= maxst)
   {
      mfft_bigpass *bp = &x->bp[x->nbp];
      if (!stage_g[cur])
      {  /* one stage, operands staged whole */
         bp->sliced = 0; bp->nops = (uint32_t)(st_off[cur + 1] - st_off[cur]); bp->nstaged = 0;
         for (k = st_off[cur]; k < st_off[cur + 1]; k++)
         {
            const mfft_op *o = &ops[k];
            const uint32_t hasB = (o->pB != MFFT_NONE && o->pB != o->pA);
            const uint32_t ns = ((o->pS == o->pA || o->pT == o->pA) ? 1u : 0u) + ((hasB && (o->pS == o->pB || o->pT == o->pB)) ? 1u : 0u);
            if (ns > bp->nstaged) bp->nstaged = ns;
         }
         bp->d_ops = (mfft_op *) mfft_upload(ops + st_off[cur], sizeof(mfft_op) * (bp->nops ? bp->nops : 1));
         if (!bp->d_ops) goto done;
         x->nbp++; cur++;
         continue;
      }
      {  /* grow a window of slice-local stages while every connected component still fits a tile */
         uint32_t e = cur, G = stage_g[cur], have = 0;
         mfft_pass best; uint32_t best_gs = 0, best_nchv = 0, best_e = 0;
         memset(&best, 0, sizeof best);
         for (;;)
         {
            uint32_t nchv = 0, gs = 0, i, maxp; size_t cb, n = st_off[e + 1] - st_off[cur]; mfft_pass cand; int r;
            for (i = 0; i < 6 && !nchv; i++)              /* the smallest slice the window's rotations allow */
               if (NCH % nchv_ok[i] == 0 && G % (NCH / nchv_ok[i]) == 0) nchv = nchv_ok[i];
            if (!nchv) break;
            gs = NCH / nchv; cb = mfft_dev_sliced_coeff_bytes(nchv);
            maxp = 4; while (maxp * 2 * cb <= budget && maxp < 256) maxp *= 2;
            for (k = 0; k < n; k++)
            {
               mfft_op *o = &win[k]; *o = ops[st_off[cur] + k];
               o->eSA = scale_e(o->eSA, gs); o->eSB = scale_e(o->eSB, gs); o->eTA = scale_e(o->eTA, gs); o->eTB = scale_e(o->eTB, gs);
            }
            r = mfft_window_pass_build(&cand, win, n, S, maxp, 128ull * nchv, last_read, live_out, NULL);
            if (r == -2) goto done;
            if (r == 0 && cand.nany)
            {  /* an op the tile executor would have to decode: not slice-safe */
               free(cand.tiles); free(cand.pos); free(cand.ops); free(cand.stoff); r = -1;
            }
            if (r != 0) break;
            if (have) { free(best.tiles); free(best.pos); free(best.ops); free(best.stoff); }
            best = cand; best_gs = gs; best_nchv = nchv; best_e = e; have = 1;
            if (e == maxst || !stage_g[e + 1]) break;
            e++; G = (uint32_t) gcd64(G, stage_g[e]);
         }
         if (!have)
         {  /* not even one stage fits as a sliced pass: run it whole */
            stage_g[cur] = 0;
            continue;
         }
         {  /* adjacent slices per CTA: contiguous R x 16-byte runs in HBM (full sectors) as long as the
               window's components still fit the tile */
            uint32_t R; size_t n = st_off[best_e + 1] - st_off[cur], cb = mfft_dev_sliced_coeff_bytes(best_nchv);
            bp->R = 1;
            for (R = 8; R >= 2; R >>= 1)
            {
               uint32_t maxp = 4; mfft_pass cand;
               if (best_gs % R || 4 * cb * R > budget) continue;
               while (maxp * 2 * cb * R <= budget && maxp < 256) maxp *= 2;
               for (k = 0; k < n; k++)
               {
                  mfft_op *o = &win[k]; *o = ops[st_off[cur] + k];
                  o->eSA = scale_e(o->eSA, best_gs); o->eSB = scale_e(o->eSB, best_gs); o->eTA = scale_e(o->eTA, best_gs); o->eTB = scale_e(o->eTB, best_gs);
               }
               if (mfft_window_pass_build(&cand, win, n, S, maxp, 128ull * best_nchv, last_read, live_out, NULL) != 0) continue;
               if (cand.nany) { free(cand.tiles); free(cand.pos); free(cand.ops); free(cand.stoff); continue; }
               free(best.tiles); free(best.pos); free(best.ops); free(best.stoff);
               best = cand; bp->R = R;
               break;
            }
         }
         bp->sliced = 1; bp->gs = best_gs; bp->nchv = best_nchv; bp->pass = best;
         bp->d.d_tiles = (mfft_tile *) mfft_upload(best.tiles, sizeof(mfft_tile) * (best.ntiles ? best.ntiles : 1));
         bp->d.d_pos = (uint32_t *) mfft_upload(best.pos, sizeof(uint32_t) * (best.npos_total ? best.npos_total : 1));
         bp->d.d_ops = (mfft_tileop *) mfft_upload(best.ops, sizeof(mfft_tileop) * (best.nops_total ? best.nops_total : 1));
         bp->d.d_stoff = (uint32_t *) mfft_upload(best.stoff, sizeof(uint32_t) * (best.nstoff ? best.nstoff : 1));
         x->nbp++;
         if (!bp->d.d_tiles || !bp->d.d_pos || !bp->d.d_ops || !bp->d.d_stoff) goto done;
         cur = best_e + 1;
      }
   }
   rc = 0;
done:
   free(ops); free(win); free(st_off); free(last_read); free(stage_g);
   return rc;
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
   if (x->bp)
   {
      for (i = 0; i < x->nbp; i++)
      {
         mfft_bigpass *b = &x->bp[i];
         mfft_dev_free(b->d.d_tiles); mfft_dev_free(b->d.d_pos); mfft_dev_free(b->d.d_ops); mfft_dev_free(b->d.d_stoff); mfft_dev_free(b->d_ops);
         free(b->pass.tiles); free(b->pass.pos); free(b->pass.ops); free(b->pass.stoff);
      }
      free(x->bp);
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
            const char *e3 = getenv("MPIRFFT_BIG_INPLACE");
            x->cs = 1; x->shift = 0;
            x->d_cw = (int32_t *) mfft_dev_alloc((size_t) 2*half_blocks*(l/2)*sizeof(int32_t));
            if (!x->d_cw) { rc = MPIRFFT_ENODEV; goto fail; }
            if (!(e3 && e3[0] == '0') && l % 64 == 0)
            {  /* in place on half 0 with multi-layer sliced passes; MPIRFFT_BIG_INPLACE=0 keeps one launch per layer */
               uint8_t *live = (uint8_t *) calloc(S, 1);
               if (!live) goto fail;
               for (k = 0; k < nout; k++) live[s->phys[k]] = 1;
               if (big_build(x, s, l, live) == 0) x->big = 1;
               else
               {  /* keep the per-layer path */
                  uint32_t q;
                  for (q = 0; q < x->nbp; q++)
                  {
                     mfft_bigpass *b = &x->bp[q];
                     mfft_dev_free(b->d.d_tiles); mfft_dev_free(b->d.d_pos); mfft_dev_free(b->d.d_ops); mfft_dev_free(b->d.d_stoff); mfft_dev_free(b->d_ops);
                     free(b->pass.tiles); free(b->pass.pos); free(b->pass.ops); free(b->pass.stoff);
                  }
                  free(x->bp); x->bp = NULL; x->nbp = 0;
               }
               free(live);
            }
         } else if (x->shift && !(e2 && e2[0] == '1') && (l % 2 == 0) && l >= 64 && NW_of(s) % 128 == 0)
         {  /* the scaling ops were emitted but the transform stays on the old path: they do the scaling there too */
            x->shift = 0;
         }
      }
      mv = (mfft_move *) malloc(sizeof(mfft_move)*(nout ? nout : 1));
      if (!mv) goto fail;
      for (k = 0; k < nout; k++) { mv[k].src_slot = x->big ? s->phys[k] : s->slot[k]; mv[k].dst_pos = dst_of[k]; }
      x->d_moves = (mfft_move *) mfft_upload(mv, sizeof(mfft_move)*(nout ? nout : 1));
      free(mv);
      if (!x->d_moves) { rc = MPIRFFT_ENODEV; goto fail; }
      if (mfft_dsched_upload(&x->ds, s) != 0) { rc = MPIRFFT_ENODEV; goto fail; }
   }
   if (getenv("MPIRFFT_VERBOSE"))
      fprintf(stderr, "mpirfft xform: l=%u S=%u batch=%u -> %s\n", l, S, nbatch,
              x->fused ? "fused tiles" : (x->big ? "carry-save, in place, sliced multi-layer passes" : (x->cs ? "carry-save stage kernel" : "ballot-carry stage kernel")));
   if (getenv("MPIRFFT_VERBOSE") && x->big)
   {
      uint32_t q;
      for (q = 0; q < x->nbp; q++)
         if (x->bp[q].sliced) fprintf(stderr, "   pass %u: sliced, stride %u chunks, %u chunks per slice, %u slices per CTA, %u tiles of <= %u positions, %u stages\n",
                                      q, x->bp[q].gs, x->bp[q].nchv, x->bp[q].R, x->bp[q].pass.ntiles, x->bp[q].pass.max_npos, x->bp[q].pass.nstages);
         else fprintf(stderr, "   pass %u: whole coefficients, %u ops\n", q, x->bp[q].nops);
   }
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
   if (x->big)
   {
      if (mfft_dev_cs_init(slab, x->d_cw, &x->g, x->d_batch, x->nbatch, stream) != 0) return MPIRFFT_ENODEV;
      for (i = 0; i < x->nbp; i++)
      {
         const mfft_bigpass *b = &x->bp[i];
         if (b->sliced)
         {
            if (mfft_dev_run_tiles_sliced(slab, x->d_cw, &x->g, b->gs, b->nchv, b->R, b->d.d_tiles, b->pass.ntiles, b->d.d_pos, b->d.d_ops,
                                          b->d.d_stoff, b->pass.max_npos, b->pass.max_nops, x->d_batch, x->nbatch, stream) != 0) return MPIRFFT_ENODEV;
         }
         else if (mfft_dev_run_stage_cs_ip(slab, x->d_cw, &x->g, b->d_ops, b->nops, x->d_batch, x->nbatch, b->nstaged, stream) != 0) return MPIRFFT_ENODEV;
      }
      if (mfft_dev_finalize_cs(dst, x->dst_stride, x->d_dst_base, slab, x->d_cw, &x->g, x->d_moves, x->nout, x->d_batch,
                               x->nbatch, x->normalise, stream) != 0) return MPIRFFT_ENODEV;
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
