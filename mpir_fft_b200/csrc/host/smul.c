/* smul.c -- new_mpn_mul with the MFA sharded over the GPUs of one box.
 *
 * The n1 columns are block-partitioned over the `world` ranks for the column passes, the live rows
 * (physical rows 0..trunc_rows-1 after the truncated column FFT) for the row passes and the
 * pointwise products; the exchange between the two is one all-to-all per transform.  After a
 * truncated forward column pass the valid rows are physically contiguous, so a rank's send
 * buffer [row][local column] is already ordered by destination rank and needs no pack kernel;
 * the receive side is unpacked into row order by one block-gather.
 *
 * This file only runs the LOCAL phases on plan-owned HBM buffers; the collectives between them
 * (all-to-all, the halo all-gather and the carry hand-off of the recombine) are issued by the
 * host language on the exchange buffers the plan exposes -- NCCL through torch.distributed in
 * mpir_fft_b200/sharded.py.  Phase order per product:
 *
 *    0 fwd_cols(i1)  A  1 fwd_rows(0)     0 fwd_cols(i2)  A  1 fwd_rows(1)     2 pointwise
 *    3 inv_rows      B  4 inv_cols        C  5 unpack  H  6 combine  carry
 *
 *    A: all_to_all  send -> recv   split by rows    (counts rows_h * ncl blocks to rank h)
 *    B: all_to_all  send -> work   split by columns (counts nrl * ncl blocks from every rank)
 *    C: same as A;  H: all_gather of the last `halo` coefficients of every rank
 */
#define _POSIX_C_SOURCE 200809L
#include "xform.h"
#include "../../../include/mpirfft_b200.h"
#include <stdlib.h>
#include <string.h>

struct mpirfft_smul_plan {
   mp_size_t n1, n2; mp_bitcnt_t depth, w;
   mpirfft_mul_params p;
   uint32_t rank, world, l, pitch;
   uint32_t ncl, c0;              /* my columns [c0, c0+ncl) */
   uint32_t r0, r1, nrl;          /* my rows [r0, r1) of the trunc_rows live rows */
   uint32_t halo;                 /* coefficients of the previous rank that reach into my limbs */
   uint64_t limb_lo, limb_hi;     /* my window of the result: limbs [limb_lo, limb_hi) */
   mfft_xform fcol, frow, irow, icol;
   limb_t *work, *send, *recv, *unp, *Z, *Y, *out;
   uint32_t *d_unpack, *d_idx, *d_carry;
   mpirfft_mulmod_plan *mm;       /* coefficient rings above 512 limbs: pointwise products through a transform */
   void *combine_work;
   size_t work_blocks, send_blocks, recv_blocks;
};

static void rows_of(uint32_t tr, uint32_t rank, uint32_t world, uint32_t *lo, uint32_t *hi)
{
   uint32_t base = tr / world, extra = tr % world;
   *lo = rank*base + (rank < extra ? rank : extra);
   *hi = *lo + base + (rank < extra ? 1 : 0);
}

void mpirfft_smul_plan_destroy(mpirfft_smul_plan *pl)
{
   if (!pl) return;
   mfft_lock();
   mfft_xform_free(&pl->fcol); mfft_xform_free(&pl->frow); mfft_xform_free(&pl->irow); mfft_xform_free(&pl->icol);
   mfft_dev_free(pl->work); mfft_dev_free(pl->send); mfft_dev_free(pl->recv); mfft_dev_free(pl->unp);
   mfft_dev_free(pl->Z); mfft_dev_free(pl->Y); mfft_dev_free(pl->out);
   mfft_dev_free(pl->d_unpack); mfft_dev_free(pl->d_idx); mfft_dev_free(pl->d_carry); mfft_dev_free(pl->combine_work);
   mfft_unlock();
   if (pl->mm) mpirfft_mulmod_plan_destroy(pl->mm);
   free(pl);
}

static uint32_t ilog2u(uint64_t x) { uint32_t b = 0; while (((uint64_t)1 << b) < x) b++; return b; }

int mpirfft_smul_plan_create(mpirfft_smul_plan **out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth,
                             mp_bitcnt_t w, int rank, int world)
{
   mpirfft_smul_plan *pl; int rc; uint64_t NW, sq, nrows2, tr, i, j, k0, k1;
   mfft_sched *s = NULL; mfft_batch *b = NULL; uint32_t *dst_of = NULL, *dst_base = NULL, *tab = NULL;
   uint32_t d1, d2, g;
   *out = NULL;
   if (world < 1 || rank < 0 || rank >= world) return MPIRFFT_EINVAL;
   pl = (mpirfft_smul_plan *) calloc(1, sizeof(*pl));
   if (!pl) return MPIRFFT_ENOMEM;
   pl->n1 = n1; pl->n2 = n2; pl->depth = depth; pl->w = w; pl->rank = (uint32_t) rank; pl->world = (uint32_t) world;
   if ((rc = mpirfft_mul_params_get(&pl->p, n1, n2, depth, w)) != 0) { free(pl); return rc; }
   sq = pl->p.sqrt; nrows2 = pl->p.n2; tr = pl->p.trunc_rows; NW = pl->p.n*w;
   if (sq % (uint64_t) world || tr < (uint64_t) world) { free(pl); return MPIRFFT_EINVAL; }
   pl->l = (uint32_t) pl->p.limbs; pl->pitch = mfft_pitch(pl->l);
   pl->ncl = (uint32_t)(sq/world); pl->c0 = pl->ncl*(uint32_t) rank;
   rows_of((uint32_t) tr, (uint32_t) rank, (uint32_t) world, &pl->r0, &pl->r1); pl->nrl = pl->r1 - pl->r0;
   d1 = ilog2u(nrows2); d2 = ilog2u(sq);
   /* recombine window: coefficients k in [k0, k1) are mine; my limbs start where coefficient k0 does */
   k0 = (uint64_t) pl->r0*sq; k1 = (uint64_t) pl->r1*sq;
   pl->limb_lo = (k0*pl->p.bits1)/64;
   pl->limb_hi = (pl->rank + 1 == pl->world) ? (uint64_t)(n1 + n2) : (k1*pl->p.bits1)/64;
   if (pl->limb_hi > (uint64_t)(n1 + n2)) pl->limb_hi = (uint64_t)(n1 + n2);
   if (pl->limb_lo > pl->limb_hi) pl->limb_lo = pl->limb_hi;
   pl->halo = pl->rank ? (uint32_t)(NW/pl->p.bits1 + 2) : 0;
   /* every rank hands its last NW/bits1 + 2 coefficients to its successor (the halo all-gather takes
      them from every rank, rank 0 and the last one included): each rank must own at least that many */
   if (NW/pl->p.bits1 + 2 > (tr/(uint64_t) world)*sq) { free(pl); return MPIRFFT_EINVAL; }

   mfft_lock();
   if ((rc = mfft_try_device()) != 0) goto fail;
   rc = MPIRFFT_ENOMEM;
   b = (mfft_batch *) calloc(sq > tr ? sq : tr, sizeof(mfft_batch));
   dst_of = (uint32_t *) malloc(sizeof(uint32_t)*(nrows2 > sq ? nrows2 : sq));
   dst_base = (uint32_t *) malloc(sizeof(uint32_t)*(sq > tr ? sq : tr));
   if (!b || !dst_of || !dst_base) goto fail;

   /* forward column transforms on my columns (mul_fft.c:2374-2389); output rows s < tr gathered to
      send[s*ncl + cl] */
   if (!(s = mfft_sched_new((uint32_t) nrows2, NW))) goto fail;
   if (mfft_sched_emit(s, MFFT_T_FFT_TRUNC, 0, 1, nrows2/2, w*sq, w, 0, 1, tr) != 0) { rc = MPIRFFT_EINVAL; goto fail; }
   mfft_sched_revbin(s, 0, 1, d1);
   for (j = 0; j < pl->ncl; j++) { b[j].base = (uint32_t) j; b[j].parity = 0; b[j].col = pl->c0 + (uint32_t) j; dst_base[j] = (uint32_t) j; }
   /* logical row revbin(s) is the s-th live row: gather it to row s */
   for (i = 0; i < tr; i++) dst_of[i] = 0;
   {
      uint32_t *of = (uint32_t *) malloc(sizeof(uint32_t)*nrows2);
      if (!of) goto fail;
      /* xform outputs are indexed by logical position k < nout; we need logical rows revbin(s).
         Relabel once more so that logical s is the s-th live row (undoing the reference's revbin
         relabel, which only exists to make the rows natural for a single address space). */
      mfft_sched_revbin(s, 0, 1, d1);
      for (i = 0; i < tr; i++) of[i] = (uint32_t) i;
      rc = mfft_xform_build(&pl->fcol, s, pl->l, pl->ncl, nrows2*pl->ncl, b, pl->ncl, of, (uint32_t) tr, dst_base, pl->ncl, 0, 0);
      free(of); s = NULL;
      if (rc != 0) goto fail;
   }
   /* forward row transforms on my rows (2392-2408): natural column order, normalised */
   rc = MPIRFFT_ENOMEM;
   if (!(s = mfft_sched_new((uint32_t) sq, NW))) goto fail;
   if (mfft_sched_emit(s, MFFT_T_FFT, 0, 1, sq/2, w*nrows2, 0, 0, 0, 0) != 0) { rc = MPIRFFT_EINVAL; goto fail; }
   mfft_sched_revbin(s, 0, 1, d2);
   for (i = 0; i < pl->nrl; i++) { b[i].base = (uint32_t)(i*sq); b[i].parity = 0; b[i].col = 0; dst_base[i] = (uint32_t)(i*sq); }
   for (j = 0; j < sq; j++) dst_of[j] = (uint32_t) j;
   rc = mfft_xform_build(&pl->frow, s, pl->l, 1, (uint64_t) pl->nrl*sq, b, pl->nrl, dst_of, (uint32_t) sq, dst_base, 1, 0, 1);
   s = NULL;
   if (rc != 0) goto fail;
   /* inverse row transforms (2942-2956); column j goes to the rank that owns it:
      send[(j/ncl)*nrl*ncl + row*ncl + j%ncl] */
   rc = MPIRFFT_ENOMEM;
   if (!(s = mfft_sched_new((uint32_t) sq, NW))) goto fail;
   mfft_sched_revbin(s, 0, 1, d2);
   if (mfft_sched_emit(s, MFFT_T_IFFT, 0, 1, sq/2, w*nrows2, 0, 0, 0, 0) != 0) { rc = MPIRFFT_EINVAL; goto fail; }
   for (i = 0; i < pl->nrl; i++) { b[i].base = (uint32_t)(i*sq); b[i].parity = 0; b[i].col = 0; dst_base[i] = (uint32_t)(i*pl->ncl); }
   for (j = 0; j < sq; j++) dst_of[j] = (uint32_t)((j/pl->ncl)*pl->nrl*pl->ncl + (j % pl->ncl));
   rc = mfft_xform_build(&pl->irow, s, pl->l, 1, (uint64_t) pl->nrl*sq, b, pl->nrl, dst_of, (uint32_t) sq, dst_base, 1, 0, 0);
   s = NULL;
   if (rc != 0) goto fail;
   /* inverse column transforms (2959-2976) + the 2^-(depth+1) scaling and normalisation of
      3256-3260; input rows are the live rows in order (logical s = s-th live row, as above) */
   rc = MPIRFFT_ENOMEM;
   if (!(s = mfft_sched_new((uint32_t) nrows2, NW))) goto fail;
   if (mfft_sched_emit(s, MFFT_T_IFFT_TRUNC, 0, 1, nrows2/2, w*sq, w, 0, 1, tr) != 0) { rc = MPIRFFT_EINVAL; goto fail; }
   for (j = 0; j < pl->ncl; j++) { b[j].base = (uint32_t) j; b[j].parity = 0; b[j].col = pl->c0 + (uint32_t) j; dst_base[j] = (uint32_t) j; }
   for (i = 0; i < tr; i++) dst_of[i] = (uint32_t) i;
   rc = mfft_xform_build(&pl->icol, s, pl->l, pl->ncl, nrows2*pl->ncl, b, pl->ncl, dst_of, (uint32_t) tr, dst_base, pl->ncl,
                         (uint32_t)(2*NW - (depth + 1)), 1);
   s = NULL;
   if (rc != 0) goto fail;

   /* buffers */
   rc = MPIRFFT_ENOMEM;
   pl->work_blocks = (size_t) 2*nrows2*pl->ncl;
   pl->send_blocks = (size_t)(tr*pl->ncl > (uint64_t) pl->nrl*sq ? tr*pl->ncl : (uint64_t) pl->nrl*sq);
   pl->recv_blocks = (size_t) pl->nrl*sq;
   {
      size_t bb = (size_t) pl->pitch*sizeof(limb_t), nout = (size_t)(pl->limb_hi - pl->limb_lo) + 2;
      pl->work = (limb_t *) mfft_dev_alloc(pl->work_blocks*bb);
      pl->send = (limb_t *) mfft_dev_alloc(pl->send_blocks*bb);
      pl->recv = (limb_t *) mfft_dev_alloc(pl->recv_blocks*bb);
      pl->unp = (limb_t *) mfft_dev_alloc((2*pl->recv_blocks + pl->halo)*bb);     /* row-pass work slab (2 halves) */
      pl->Z = (limb_t *) mfft_dev_alloc(2*pl->recv_blocks*bb);
      pl->Y = (limb_t *) mfft_dev_alloc(pl->recv_blocks*bb);
      pl->out = (limb_t *) mfft_dev_alloc(nout*sizeof(limb_t));
      pl->combine_work = mfft_dev_alloc(mfft_dev_combine_work(nout));
      pl->d_carry = (uint32_t *) mfft_dev_alloc(64);
      if (!pl->work || !pl->send || !pl->recv || !pl->unp || !pl->Z || !pl->Y || !pl->out || !pl->combine_work || !pl->d_carry) goto fail;
   }
   /* unpack table: recv holds, for every source rank g, [my row][g's column] */
   tab = (uint32_t *) malloc(sizeof(uint32_t)*pl->recv_blocks);
   if (!tab) goto fail;
   for (i = 0; i < pl->nrl; i++)
      for (j = 0; j < sq; j++)
      {
         g = (uint32_t)(j/pl->ncl);
         tab[i*sq + j] = (uint32_t)((uint64_t) g*pl->nrl*pl->ncl + i*pl->ncl + (j % pl->ncl));
      }
   pl->d_unpack = (uint32_t *) mfft_upload(tab, sizeof(uint32_t)*pl->recv_blocks);
   for (i = 0; i < pl->recv_blocks; i++) tab[i] = (uint32_t) i;
   pl->d_idx = (uint32_t *) mfft_upload(tab, sizeof(uint32_t)*pl->recv_blocks);
   if (!pl->d_unpack || !pl->d_idx) { rc = MPIRFFT_ENODEV; goto fail; }
   free(tab); free(b); free(dst_of); free(dst_base);
   mfft_unlock();
   if (pl->l > 512)
   {  /* no warp-level product kernel for such coefficients: recurse like the reference does for its
         pointwise products (fft_mulmod_2expp1, mul_fft.c:3125) -- one batched transform-based mulmod
         over all the rank's coefficients */
      mp_bitcnt_t d2, w2;
      if (mpirfft_mulmod_params((mp_size_t) pl->l, &d2, &w2) != 0 ||
          mpirfft_mulmod_plan_create(&pl->mm, (mp_size_t) pl->l, d2, w2, pl->recv_blocks) != 0)
      { mpirfft_smul_plan_destroy(pl); return MPIRFFT_EINVAL; }
   }
   *out = pl;
   return 0;
fail:
   if (s) mfft_sched_free(s);
   free(tab); free(b); free(dst_of); free(dst_base);
   mfft_unlock();
   mpirfft_smul_plan_destroy(pl);
   return rc;
}

/* geometry the host language needs to issue the collectives (all counts in limbs) */
int mpirfft_smul_info(const mpirfft_smul_plan *pl, mpirfft_smul_layout *o)
{
   memset(o, 0, sizeof(*o));
   o->block_limbs = pl->pitch; o->ncl = pl->ncl; o->nrl = pl->nrl; o->r0 = pl->r0; o->trunc_rows = (uint32_t) pl->p.trunc_rows;
   o->n1cols = (uint32_t) pl->p.sqrt; o->halo = pl->halo; o->halo_send = (uint32_t)((pl->p.n*pl->w)/pl->p.bits1 + 2); o->limb_lo = pl->limb_lo; o->limb_hi = pl->limb_hi;
   o->send = pl->send; o->recv = pl->recv; o->work = pl->work; o->unp = pl->unp; o->out = pl->out;
   o->send_limbs = pl->send_blocks*pl->pitch; o->recv_limbs = pl->recv_blocks*pl->pitch;
   o->work_limbs = pl->work_blocks*pl->pitch;
   return 0;
}

int mpirfft_smul_phase(mpirfft_smul_plan *pl, int phase, int which, const mp_limb_t *d_in, void *stream)
{
   const mpirfft_mul_params *p = &pl->p;
   uint64_t sq = p->sqrt, tr = p->trunc_rows;
   switch (phase)
   {
   case 0:   /* split my columns, column FFTs, rows gathered (by destination rank) into send */
      if (mfft_dev_split_cols(pl->work, pl->l, pl->pitch, (const limb_t *) d_in, (uint64_t)(which ? pl->n2 : pl->n1), p->bits1,
                              which ? p->j2 : p->j1, tr, pl->ncl, (uint32_t) sq, pl->c0, stream)) return MPIRFFT_ENODEV;
      return mfft_xform_exec(&pl->fcol, pl->work, pl->send, stream);
   case 1:   /* unpack the received [source][row][col] into rows, row FFTs -> Z (which 0) or Y (which 1) */
      if (mfft_dev_gather_blocks(pl->unp, pl->recv, pl->d_unpack, pl->recv_blocks, pl->pitch, stream)) return MPIRFFT_ENODEV;
      return mfft_xform_exec(&pl->frow, pl->unp, which ? pl->Y : pl->Z, stream);
   case 2:
      if (pl->mm) return mpirfft_mulmod_plan_exec(pl->mm, (mp_limb_t *) pl->Z, (const mp_limb_t *) pl->Z, (const mp_limb_t *) pl->Y, pl->pitch, -1, stream);
      if (mfft_dev_pointwise(pl->Z, pl->Y, pl->d_idx, (uint32_t) pl->recv_blocks, pl->l, pl->pitch, stream)) return MPIRFFT_ENODEV;
      return 0;
   case 3:   /* row IFFTs, columns gathered by owner into send */
      return mfft_xform_exec(&pl->irow, pl->Z, pl->send, stream);
   case 4:   /* (exchange B delivered [row][my col] into work) column IFFTs, scaled + normalised -> send */
      return mfft_xform_exec(&pl->icol, pl->work, pl->send, stream);
   case 5:   /* (exchange C) unpack into rows behind the halo slots */
      if (mfft_dev_gather_blocks(pl->unp + (size_t) pl->halo*pl->pitch, pl->recv, pl->d_unpack, pl->recv_blocks, pl->pitch, stream)) return MPIRFFT_ENODEV;
      return 0;
   case 6:   /* (halo filled) my window of the result, carry-in 0; carry out of the window -> d_carry */
   {
      uint64_t k0 = (uint64_t) pl->r0*sq, nco = pl->recv_blocks, total = pl->limb_hi - pl->limb_lo;
      uint64_t last = p->j1 + p->j2 - 1;                          /* coefficients >= last are not part of the product */
      uint64_t first_local = k0 - pl->halo;                        /* global index of local coefficient 0 */
      uint64_t ncoef = (first_local + pl->halo + nco > last) ? (last > first_local ? last - first_local : 0) : pl->halo + nco;
      uint64_t base_bit = pl->limb_lo*64 - first_local*p->bits1;
      if (!total) { if (mfft_dev_memset0(pl->d_carry, 4, stream)) return MPIRFFT_ENODEV; return 0; }
      if (mfft_dev_combine_window(pl->out, total, pl->unp, pl->l, pl->pitch, p->bits1, ncoef, base_bit, pl->d_carry, pl->combine_work, stream)) return MPIRFFT_ENODEV;
      return 0;
   }
   default: return MPIRFFT_EINVAL;
   }
}

/* add the carry handed over by the previous rank into my window; *carry_out = what I hand on:
 * the carry that left my window in phase 6 plus the ripple of carry_in (mutually exclusive in
 * practice, both tiny) */
int mpirfft_smul_carry(mpirfft_smul_plan *pl, unsigned carry_in, unsigned *carry_out, void *stream)
{
   uint64_t total = pl->limb_hi - pl->limb_lo; uint32_t own = 0, rip = 0;
   if (mfft_dev_d2h(&own, pl->d_carry, 4, stream) || mfft_dev_sync(stream)) return MPIRFFT_ENODEV;
   if (!total) { *carry_out = carry_in + own; return 0; }
   if (carry_in)
   {
      if (mfft_dev_add_small(pl->out, total, carry_in, pl->d_carry + 1, stream)) return MPIRFFT_ENODEV;
      if (mfft_dev_d2h(&rip, pl->d_carry + 1, 4, stream) || mfft_dev_sync(stream)) return MPIRFFT_ENODEV;
   }
   *carry_out = own + rip;
   return 0;
}

/* the last `count` coefficients I own (what the next rank needs as its halo) / my halo slots */
mp_limb_t *mpirfft_smul_tail_blocks(mpirfft_smul_plan *pl, unsigned count)
{ if ((size_t) count > pl->recv_blocks) return NULL;
  return (mp_limb_t *)(pl->unp + ((size_t) pl->halo + pl->recv_blocks - count)*pl->pitch); }
