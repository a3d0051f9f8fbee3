/* tile.c -- partition a schedule into fused passes for the shared-memory tile executor.
 *
 * The reference's MFA exists for cache locality (README:74-91): sub-transforms of ~sqrt(2n)
 * coefficients fit the cache.  On B200 the same idea one level down: a window of consecutive
 * stages is cut into connected components of positions (the radix-2^k sub-transforms, or the
 * irregular pieces of the truncated variants); components are packed into tiles of at most
 * `max_npos` coefficients, and one CTA runs ALL the stages of a tile inside shared memory, so a
 * pass costs one read and one write of the live slab instead of one per radix-2 layer.
 *
 * The input is the position-based view of the schedule (mfft_op.pA/pB/pS/pT/pstage): ops work in
 * place on physical blocks of slab half 0, so passes need no ping-pong.
 */
#include "tile.h"
#include <stdlib.h>
#include <string.h>

static uint32_t uf_find(uint32_t *par, uint32_t x)
{
   while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; }
   return x;
}

/* union; returns the size of the merged component */
static uint32_t uf_union(uint32_t *par, uint32_t *sz, uint32_t a, uint32_t b)
{
   a = uf_find(par, a); b = uf_find(par, b);
   if (a == b) return sz[a];
   if (sz[a] < sz[b]) { uint32_t t = a; a = b; b = t; }
   par[b] = a; sz[a] += sz[b];
   return sz[a];
}

static int cmp_pstage(const void *x, const void *y)
{
   const mfft_op *a = (const mfft_op *) x, *b = (const mfft_op *) y;
   if (a->pstage != b->pstage) return a->pstage < b->pstage ? -1 : 1;
   return 0;
}

void mfft_passes_free(mfft_passes *P)
{
   uint32_t i;
   if (!P) return;
   for (i = 0; i < P->npasses; i++)
   {
      free(P->pass[i].tiles); free(P->pass[i].pos); free(P->pass[i].ops); free(P->pass[i].stoff);
   }
   free(P->pass);
   memset(P, 0, sizeof(*P));
}

/* Recognise the common aligned shapes (see MFFT_K_* in mfft_internal.h).  NW = ring bits. */
static int fold_term(int sign, uint32_t e, uint64_t NW, uint32_t *yc, uint32_t *neg)
{
   if (e >= NW) { e -= (uint32_t) NW; sign = -sign; }
   if (e % 128) return -1;
   *yc = e / 128; *neg = (sign < 0);
   return 0;
}

static void classify_op(mfft_tileop *d, uint64_t NW)
{
   const int hasB = (d->b != 0xFFFF), hasT = (d->t != 0xFFFF);
   uint32_t ySA = 0, nSA = 0, ySB = 0, nSB = 0, yTA = 0, nTA = 0, yTB = 0, nTB = 0;
   d->kind = MFFT_K_ANY; d->kparam = 0;
   if (d->cSA | d->cSB | d->cTA | d->cTB) return;
   if (NW % 128 || !d->sSA) return;
   /* the two bit-granular shapes of the truncated inverse that stay chunk-local in carry-save form */
   if (!hasB && !hasT && d->sSA == 1 && d->eSA == 1) { d->kind = MFFT_K_DBL; return; }
   if (!hasB && !hasT && d->sSA == 1 && d->eSA > 2*NW - 32 && d->eSA < 2*NW)
   { d->kind = MFFT_K_SHR; d->kparam = (uint32_t)(2*NW - d->eSA); return; }
   if (hasB && !hasT && d->sSA == 1 && d->sSB == 1 && d->eSA == 2*NW - 1 && d->eSB == 2*NW - 1) { d->kind = MFFT_K_HALF; return; }
   if (hasB && d->sSA == 1 && d->eSA == 1 && d->sSB == -1 && d->eSB == 0)
   {  /* S = 2A - B, optionally T = +-(A - B) rotated by whole chunks (1630-1631, 1639-1647) */
      if (!hasT) { d->kind = MFFT_K_2AMB; return; }
      if (d->sTA && d->sTB && !fold_term(d->sTA, d->eTA, NW, &yTA, &nTA) && !fold_term(d->sTB, d->eTB, NW, &yTB, &nTB) &&
          yTA == yTB && nTA != nTB) { d->kind = MFFT_K_2AMB; d->kparam = yTA | (nTA << 31); }
      return;
   }
   if (fold_term(d->sSA, d->eSA, NW, &ySA, &nSA)) return;
   if (hasB) { if (!d->sSB || fold_term(d->sSB, d->eSB, NW, &ySB, &nSB)) return; }
   if (hasT)
   {
      if (!d->sTA || fold_term(d->sTA, d->eTA, NW, &yTA, &nTA)) return;
      if (hasB) { if (!d->sTB || fold_term(d->sTB, d->eTB, NW, &yTB, &nTB)) return; }
   }
   if (hasB && hasT)
   {
      if (ySA == 0 && !nSA && ySB == 0 && !nSB && yTA == yTB && nTA != nTB)
      { d->kind = MFFT_K_FWD; d->kparam = yTA | (nTA << 31); return; }
      if (ySA == 0 && !nSA && yTA == 0 && !nTA && ySB == yTB && nSB != nTB)
      { d->kind = MFFT_K_INV; d->kparam = ySB | (nSB << 31); return; }
   } else if (!hasB && !hasT)
   { d->kind = MFFT_K_ROT; d->kparam = ySA | (nSA << 31); return; }
   else if (hasB && !hasT && ySA == 0 && !nSA && ySB == 0 && !nSB)
   { d->kind = MFFT_K_ADD; return; }
}

/* Fuse pairs of consecutive layers into radix-4 units (MFFT_K_FWD4 / MFFT_K_INV4) where four
 * in-place butterflies of the same kind form a 2x2 network whose quarter-turn twiddles differ by
 * exactly NW/2 -- then the unit is lane-local in the kernel (mfft_tiles.h).  Works on one tile's op
 * slice (sorted by lstage); absorbed ops become MFFT_K_NOP and are compacted away by the caller. */
static void fuse_radix4(mfft_tileop *ops, uint32_t nops, uint32_t nstages, uint32_t npos, uint64_t NW, uint32_t *scratch)
{
   const uint32_t NCH = (uint32_t)(NW/128), NT = NCH/32;
   uint32_t *asA0 = scratch, *asA1 = scratch + npos, s, k;      /* position -> op (as operand a) per stage */
   const char *env = getenv("MPIRFFT_NO_RADIX4");
   if ((env && env[0] == '1') || NW % 128 || NCH % 32 || NT > 4 || (NT & 1)) return;
   for (s = 0; s + 1 < nstages; s += 2)
   {
      uint32_t lo0 = nops, hi0 = 0, lo1 = nops, hi1 = 0;
      for (k = 0; k < nops; k++)
      {
         if (ops[k].lstage == s) { if (k < lo0) lo0 = k; hi0 = k + 1; }
         if (ops[k].lstage == s + 1) { if (k < lo1) lo1 = k; hi1 = k + 1; }
      }
      if (lo0 >= hi0 || lo1 >= hi1) continue;
      for (k = 0; k < npos; k++) { asA0[k] = MFFT_NONE; asA1[k] = MFFT_NONE; }
      for (k = lo0; k < hi0; k++) if (ops[k].kind == MFFT_K_FWD || ops[k].kind == MFFT_K_INV)
         if (ops[k].s == ops[k].a && ops[k].t == ops[k].b) asA0[ops[k].a] = k;
      for (k = lo1; k < hi1; k++) if (ops[k].kind == MFFT_K_FWD || ops[k].kind == MFFT_K_INV)
         if (ops[k].s == ops[k].a && ops[k].t == ops[k].b) asA1[ops[k].a] = k;
      for (k = lo0; k < hi0; k++)
      {  /* layer s: X = (p0,x), X' = (y,p3); layer s+1: Y = (p0,y), Z = (x,p3).  An op with zero
            rotation (S = A+B, T = A-B) fits both the forward and the inverse shape. */
#define FWDLIKE(o) ((o)->kind == MFFT_K_FWD || ((o)->kind == MFFT_K_INV && (o)->kparam == 0))
#define INVLIKE(o) ((o)->kind == MFFT_K_INV || ((o)->kind == MFFT_K_FWD && (o)->kparam == 0))
#define YC(o) ((o)->kparam & 0x7fffffffu)
         mfft_tileop *X = &ops[k], *Xp, *Y, *Z;
         uint32_t p0, x, y, p3; int fwd, inv;
         if ((X->kind != MFFT_K_FWD && X->kind != MFFT_K_INV) || asA0[X->a] != k) continue;
         p0 = X->a; x = X->b;
         if (asA1[p0] == MFFT_NONE) continue;
         Y = &ops[asA1[p0]]; y = Y->b;
         if (y == x || asA0[y] == MFFT_NONE) continue;
         Xp = &ops[asA0[y]]; p3 = Xp->b;
         if (Xp == X || asA1[x] == MFFT_NONE) continue;
         Z = &ops[asA1[x]];
         if (Z->b != p3) continue;
         fwd = FWDLIKE(X) && FWDLIKE(Xp) && FWDLIKE(Y) && FWDLIKE(Z) && ((YC(Xp) + NCH - YC(X)) % NCH) == NCH/2;
         inv = INVLIKE(X) && INVLIKE(Xp) && INVLIKE(Y) && INVLIKE(Z) && ((YC(Z) + NCH - YC(Y)) % NCH) == NCH/2;
         if (!fwd && !inv) continue;
         X->eSA = X->kparam; X->eSB = Xp->kparam; X->eTA = Y->kparam; X->eTB = Z->kparam;
         asA0[p0] = MFFT_NONE; asA0[y] = MFFT_NONE; asA1[p0] = MFFT_NONE; asA1[x] = MFFT_NONE;
         Xp->kind = MFFT_K_NOP; Y->kind = MFFT_K_NOP; Z->kind = MFFT_K_NOP;
         if (fwd) { X->kind = MFFT_K_FWD4; X->a = (uint16_t) p0; X->b = (uint16_t) y; X->s = (uint16_t) x; X->t = (uint16_t) p3; }
         else     { X->kind = MFFT_K_INV4; X->a = (uint16_t) p0; X->b = (uint16_t) x; X->s = (uint16_t) y; X->t = (uint16_t) p3; }
#undef FWDLIKE
#undef INVLIKE
#undef YC
      }
   }
}

/* build one pass from ops[lo..hi) (sorted by pstage; window = stages s0..s1) */
static int build_pass(mfft_pass *out, const mfft_op *ops, size_t lo, size_t hi, uint32_t S,
                      uint32_t s0, uint32_t max_npos, const uint32_t *par_in, uint32_t *scratch,
                      const uint8_t *must_store, uint64_t NW, uint32_t s1, const uint32_t *last_read,
                      const uint8_t *live_out)
{
   /* scratch: 6*S uint32: root->tile map, first-access kind, written flag, local index, order */
   uint32_t *root_tile = scratch, *first = scratch + S, *written = scratch + 2*S, *local = scratch + 3*S;
   uint32_t *tile_fill = scratch + 4*S, *par = scratch + 5*S;
   uint32_t ntiles = 0, p; size_t k;
   uint32_t *tile_npos, *tile_nops, *pos_cursor, *op_cursor;

   memcpy(par, par_in, sizeof(uint32_t) * S);
   for (p = 0; p < S; p++) { root_tile[p] = MFFT_NONE; first[p] = 0; written[p] = 0; local[p] = MFFT_NONE; }

   /* first access kind per position: 1 = read first (needs load), 2 = written first */
   for (k = lo; k < hi; k++)
   {
      const mfft_op *o = &ops[k];
      if (!first[o->pA]) first[o->pA] = 1;
      if (o->pB != MFFT_NONE && !first[o->pB]) first[o->pB] = 1;
      if (!first[o->pS]) first[o->pS] = 2;
      written[o->pS] = 1;
      if (o->pT != MFFT_NONE) { if (!first[o->pT]) first[o->pT] = 2; written[o->pT] = 1; }
   }
   if (must_store)
      for (p = 0; p < S; p++)
         if (must_store[p]) { if (!first[p]) first[p] = 1; written[p] = 1; }
   /* component sizes */
   memset(tile_fill, 0, sizeof(uint32_t) * S);
   for (p = 0; p < S; p++) if (first[p]) tile_fill[uf_find(par, p)]++;      /* size per root */
   /* pack components into tiles, first fit in order of appearance */
   {
      uint32_t cur_fill = 0, cur = MFFT_NONE;
      for (p = 0; p < S; p++)
      {
         uint32_t r;
         if (!first[p]) continue;
         r = uf_find(par, p);
         if (root_tile[r] != MFFT_NONE) continue;
         if (tile_fill[r] > max_npos) return -1;
         if (cur == MFFT_NONE || cur_fill + tile_fill[r] > max_npos) { cur = ntiles++; cur_fill = 0; }
         root_tile[r] = cur; cur_fill += tile_fill[r];
      }
   }
   out->ntiles = ntiles; out->max_npos = 0; out->max_nops = 0; out->nstages = 0; out->nany = 0; out->nr4 = 0;
   out->tiles = (mfft_tile *) calloc(ntiles ? ntiles : 1, sizeof(mfft_tile));
   tile_npos = (uint32_t *) calloc(4 * (size_t)(ntiles ? ntiles : 1), sizeof(uint32_t));
   if (!out->tiles || !tile_npos) { free(tile_npos); return -1; }
   tile_nops = tile_npos + ntiles; pos_cursor = tile_nops + ntiles; op_cursor = pos_cursor + ntiles;
   for (p = 0; p < S; p++) if (first[p]) tile_npos[root_tile[uf_find(par, p)]]++;
   for (k = lo; k < hi; k++) tile_nops[root_tile[uf_find(par, ops[k].pA)]]++;
   {
      uint32_t po = 0, oo = 0, t;
      for (t = 0; t < ntiles; t++)
      {
         out->tiles[t].pos_off = po; out->tiles[t].npos = tile_npos[t];
         out->tiles[t].op_off = oo; out->tiles[t].nops = tile_nops[t];
         pos_cursor[t] = po; op_cursor[t] = oo;
         po += tile_npos[t]; oo += tile_nops[t];
         if (tile_npos[t] > out->max_npos) out->max_npos = tile_npos[t];
         if (tile_nops[t] > out->max_nops) out->max_nops = tile_nops[t];
      }
      out->npos_total = po; out->nops_total = oo;
   }
   out->pos = (uint32_t *) malloc(sizeof(uint32_t) * (out->npos_total ? out->npos_total : 1));
   out->ops = (mfft_tileop *) calloc(out->nops_total ? out->nops_total : 1, sizeof(mfft_tileop));
   if (!out->pos || !out->ops) { free(tile_npos); return -1; }
   for (p = 0; p < S; p++)
   {
      uint32_t t;
      if (!first[p]) continue;
      t = root_tile[uf_find(par, p)];
      local[p] = pos_cursor[t] - out->tiles[t].pos_off;
      {  /* a value nobody reads after this pass and that is not an output of the schedule is dead:
            the truncated transforms leave many of those behind (rows they synthesise and fold away) */
         const int live = (last_read[p] > s1) || !live_out || live_out[p] || (must_store && must_store[p]);
         out->pos[pos_cursor[t]++] = p | (first[p] == 1 ? MFFT_TILE_LOAD : 0) | ((written[p] && live) ? MFFT_TILE_STORE : 0);
      }
   }
   for (k = lo; k < hi; k++)      /* ops are in pstage order, so each tile's list is too */
   {
      const mfft_op *o = &ops[k];
      uint32_t t = root_tile[uf_find(par, o->pA)];
      mfft_tileop *d = &out->ops[op_cursor[t]++];
      d->a = (uint16_t) local[o->pA];
      d->b = (o->pB == MFFT_NONE) ? 0xFFFF : (uint16_t) local[o->pB];
      d->s = (uint16_t) local[o->pS];
      d->t = (o->pT == MFFT_NONE) ? 0xFFFF : (uint16_t) local[o->pT];
      d->eSA = o->eSA; d->eSB = o->eSB; d->eTA = o->eTA; d->eTB = o->eTB;
      d->cSA = o->cSA; d->cSB = o->cSB; d->cTA = o->cTA; d->cTB = o->cTB;
      d->sSA = o->sSA; d->sSB = o->sSB; d->sTA = o->sTA; d->sTB = o->sTB;
      d->lstage = o->pstage - s0;
      classify_op(d, NW);
      if (d->kind == MFFT_K_ANY) out->nany++;
      if (d->lstage + 1 > out->tiles[t].nstages) out->tiles[t].nstages = d->lstage + 1;
      if (d->lstage + 1 > out->nstages) out->nstages = d->lstage + 1;
   }
   free(tile_npos);
   {  /* radix-4 fusion per tile; drop the absorbed ops and the stages that became empty */
      uint32_t t; uint32_t *fs = (uint32_t *) malloc(sizeof(uint32_t) * 2 * (size_t)(out->max_npos ? out->max_npos : 1));
      if (!fs) return -1;
      out->nstages = 0; out->max_nops = 0;
      for (t = 0; t < ntiles; t++)
      {
         mfft_tile *tl = &out->tiles[t]; mfft_tileop *o = out->ops + tl->op_off;
         uint32_t k, w = 0, remap[64], used[64], ns = 0;
         if (tl->nstages > 62) { free(fs); return -1; }
         fuse_radix4(o, tl->nops, tl->nstages, tl->npos, NW, fs);
         memset(used, 0, sizeof used);
         for (k = 0; k < tl->nops; k++) if (o[k].kind != MFFT_K_NOP) used[o[k].lstage] = 1;
         for (k = 0; k < tl->nstages; k++) { remap[k] = ns; ns += used[k]; }
         for (k = 0; k < tl->nops; k++) if (o[k].kind != MFFT_K_NOP) { o[w] = o[k]; o[w].lstage = remap[o[k].lstage]; w++; }
         for (k = 0; k < w; k++) if (o[k].kind == MFFT_K_FWD4 || o[k].kind == MFFT_K_INV4) out->nr4++;
         tl->nops = w; tl->nstages = ns;
         if (ns > out->nstages) out->nstages = ns;
         if (w > out->max_nops) out->max_nops = w;
      }
      free(fs);
   }
   {  /* stage offsets per tile */
      uint32_t t, total = 0, cur = 0;
      for (t = 0; t < ntiles; t++) total += out->tiles[t].nstages + 1;
      out->stoff = (uint32_t *) malloc(sizeof(uint32_t) * (total ? total : 1));
      if (!out->stoff) return -1;
      out->nstoff = total;
      for (t = 0; t < ntiles; t++)
      {
         mfft_tile *tl = &out->tiles[t]; uint32_t st, o = 0;
         if (tl->nstages > 62) return -1;
         tl->pad = cur;
         for (st = 0; st <= tl->nstages; st++)
         {
            while (o < tl->nops && out->ops[tl->op_off + o].lstage < st) o++;
            out->stoff[cur++] = o;
         }
      }
   }
   return 0;
}

int mfft_passes_build(mfft_passes *P, const mfft_sched *s, uint32_t max_npos, const uint8_t *must_store, const uint8_t *live_out)
{
   uint32_t S = s->S, maxst = s->npstages, s0, p;
   mfft_op *ops; uint32_t *par, *sz, *par2, *sz2, *scratch, *last_read; size_t *st_off; size_t k;
   int rc = -1;
   memset(P, 0, sizeof(*P));
   if (max_npos < 4) return -1;
   last_read = (uint32_t *) calloc(S ? S : 1, sizeof(uint32_t));       /* latest pstage that reads a position */
   if (!last_read) return -1;
   for (k = 0; k < s->nops; k++)
   {
      const mfft_op *o = &s->ops[k];
      if (o->pstage > last_read[o->pA]) last_read[o->pA] = o->pstage;
      if (o->pB != MFFT_NONE && o->pstage > last_read[o->pB]) last_read[o->pB] = o->pstage;
   }
   ops = (mfft_op *) malloc(sizeof(mfft_op) * (s->nops ? s->nops : 1));
   par = (uint32_t *) malloc(sizeof(uint32_t) * 4 * (size_t) S);
   scratch = (uint32_t *) malloc(sizeof(uint32_t) * 6 * (size_t) S);
   st_off = (size_t *) calloc((size_t) maxst + 2, sizeof(size_t));
   P->pass = (mfft_pass *) calloc((size_t) maxst + 1, sizeof(mfft_pass));
   if (!ops || !par || !scratch || !st_off || !P->pass) goto done;
   sz = par + S; par2 = sz + S; sz2 = par2 + S;
   memcpy(ops, s->ops, sizeof(mfft_op) * s->nops);
   qsort(ops, s->nops, sizeof(mfft_op), cmp_pstage);       /* stable enough: same-stage ops are independent */
   for (k = 0; k < s->nops; k++) st_off[ops[k].pstage]++;
   {  /* st_off[t] = first op of stage t (1-based stages) */
      size_t acc = 0; uint32_t t;
      for (t = 0; t <= maxst + 1; t++) { size_t c = st_off[t]; st_off[t] = acc; acc += c; }
   }
   s0 = 1;
   while (s0 <= maxst)
   {
      uint32_t s1 = s0, worst = 0;
      for (p = 0; p < S; p++) { par[p] = p; sz[p] = 1; }
      for (;;)
      {  /* try to take stage s1 into the window */
         memcpy(par2, par, sizeof(uint32_t) * S); memcpy(sz2, sz, sizeof(uint32_t) * S);
         worst = 0;
         for (k = st_off[s1]; k < st_off[s1 + 1]; k++)
         {
            const mfft_op *o = &ops[k]; uint32_t m = 1;
            if (o->pB != MFFT_NONE) m = uf_union(par2, sz2, o->pA, o->pB);
            m = uf_union(par2, sz2, o->pA, o->pS);
            if (o->pT != MFFT_NONE) m = uf_union(par2, sz2, o->pA, o->pT);
            if (m > worst) worst = m;
         }
         if (worst > max_npos)
         {
            if (s1 == s0) goto done;          /* a single stage must always fit (<= 4 positions) */
            s1--; break;
         }
         memcpy(par, par2, sizeof(uint32_t) * S); memcpy(sz, sz2, sizeof(uint32_t) * S);
         if (s1 == maxst) break;
         s1++;
      }
      if (build_pass(&P->pass[P->npasses], ops, st_off[s0], st_off[s1 + 1], S, s0, max_npos, par, scratch,
                     (s1 == maxst) ? must_store : NULL, s->NW, s1, last_read, live_out) != 0) goto done;
      P->npasses++;
      s0 = s1 + 1;
   }
   rc = 0;
done:
   free(ops); free(par); free(scratch); free(st_off); free(last_read);
   if (rc != 0) mfft_passes_free(P);
   return rc;
}


/* One pass from an explicit window of ops (sorted by pstage, all of them inside the window): used by
 * the big-ring executor (bigpass.c), which cuts its own windows.  NW is the ring the ops' exponents
 * refer to (a slice of a big coefficient is a small ring of its own).  Returns -1 if a connected
 * component exceeds max_npos (the caller then shrinks the window), -2 on allocation failure. */
int mfft_window_pass_build(mfft_pass *out, const mfft_op *ops, size_t nops, uint32_t S, uint32_t max_npos,
                           uint64_t NW, const uint32_t *last_read, const uint8_t *live_out, const uint8_t *must_store)
{
   uint32_t *par, *sz, *scratch, p, s0, s1; size_t k; int rc = -2;
   memset(out, 0, sizeof(*out));
   if (!nops) return -2;
   par = (uint32_t *) malloc(sizeof(uint32_t) * 2 * (size_t) S);
   scratch = (uint32_t *) malloc(sizeof(uint32_t) * 6 * (size_t) S);
   if (!par || !scratch) goto done;
   sz = par + S;
   for (p = 0; p < S; p++) { par[p] = p; sz[p] = 1; }
   s0 = ops[0].pstage; s1 = ops[nops - 1].pstage;
   for (k = 0; k < nops; k++)
   {
      const mfft_op *o = &ops[k]; uint32_t m = 1;
      if (o->pB != MFFT_NONE) m = uf_union(par, sz, o->pA, o->pB);
      m = uf_union(par, sz, o->pA, o->pS);
      if (o->pT != MFFT_NONE) m = uf_union(par, sz, o->pA, o->pT);
      if (m > max_npos) { rc = -1; goto done; }
   }
   rc = build_pass(out, ops, 0, nops, S, s0, max_npos, par, scratch, must_store, NW, s1, last_read, live_out) == 0 ? 0 : -2;
   if (rc != 0) { free(out->tiles); free(out->pos); free(out->ops); free(out->stoff); memset(out, 0, sizeof(*out)); }
done:
   free(par); free(scratch);
   return rc;
}
