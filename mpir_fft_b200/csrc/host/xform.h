/* xform.h -- one schedule applied to a batch of strided block families, ending in a gather.
 * The building block of the sharded (multi-GPU) multiplication: a rank runs column transforms on
 * its column shard and row transforms on its row shard, each gathering into an exchange buffer. */
#ifndef MFFT_XFORM_H
#define MFFT_XFORM_H
#include "runtime.h"
#ifdef __cplusplus
extern "C" {
#endif

/* one pass of the big-ring executor (rings above 512 limbs, in place on slab half 0) */
typedef struct {
   int sliced;                                    /* 1: multi-layer pass on chunk slices; 0: one stage, operands staged whole */
   uint32_t gs, nchv, R;                          /* sliced: slice stride in chunks, chunks per slice, adjacent slices per CTA */
   mfft_pass pass; struct mfft_dpass d;           /* sliced: tile descriptors (virtual ring of nchv chunks) */
   mfft_op *d_ops; uint32_t nops, nstaged;        /* whole: the stage's ops; operands an op stages at most (outputs that overwrite inputs) */
} mfft_bigpass;

typedef struct {
   mfft_geom g; int fused; uint32_t S, nbatch, nout, dst_stride, shift; int normalise;
   mfft_dsched ds;                                /* stagewise: uploaded schedule (owns s) */
   mfft_sched *s;                                 /* fused: host schedule (owned) */
   mfft_passes P; struct mfft_dpass *dp;          /* fused passes */
   mfft_batch *d_batch, *h_batch; uint32_t *d_dst_base, *d_dstpos; mfft_move *d_moves;
   int cs; int32_t *d_cw;                         /* stagewise, carry-save kernel: carry words of both slab halves */
   int big; mfft_bigpass *bp; uint32_t nbp;       /* carry-save, in place, multi-layer sliced passes (xform.c: big_build) */
} mfft_xform;

/* s: emitted (and relabelled) schedule over S positions; ownership passes to the xform.
 * batch[nbatch]: strided families; dst_of[k], k < nout: destination position value of logical
 * output k (the gather writes block dst_base[b] + dst_of[k]*dst_stride); shift: final factor
 * 2^shift (bit exponent mod 2NW) on every output; normalise: canonical outputs.
 * slot_stride / half_blocks: geometry of the work slab (block = half*half_blocks + base_b + pos*slot_stride). */
int  mfft_xform_build(mfft_xform *x, mfft_sched *s, uint32_t l, uint32_t slot_stride, uint64_t half_blocks,
                      const mfft_batch *batch, uint32_t nbatch, const uint32_t *dst_of, uint32_t nout,
                      const uint32_t *dst_base, uint32_t dst_stride, uint32_t shift, int normalise);
int  mfft_xform_exec(const mfft_xform *x, limb_t *slab, limb_t *dst, void *stream);
void mfft_xform_free(mfft_xform *x);
int  mfft_xform_halves(const mfft_xform *x);      /* 1 (fused, in place) or 2 (stagewise ping-pong) */

#ifdef __cplusplus
}
#endif
#endif
