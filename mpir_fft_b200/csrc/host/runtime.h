/* runtime.h -- host-side plumbing shared by the C files of the library (plain C99). */
#ifndef MFFT_RUNTIME_H
#define MFFT_RUNTIME_H

#include "../mfft_internal.h"
#include "sched.h"
#include "tile.h"

#ifdef __cplusplus
extern "C" {
#endif

/* abort with a diagnostic: the reference's entry points have no error channel */
void mfft_die(const char *fn, const char *fmt, ...);
/* lazily bind the process to a device (MPIRFFT_DEVICE, default 0); dies if there is none */
void mfft_require_device(const char *fn);
int  mfft_try_device(void);            /* same, returning <0 instead of dying */
void mfft_lock(void);
void mfft_unlock(void);

/* A schedule uploaded to the device */
typedef struct {
   mfft_sched *s;          /* host copy (owned) */
   mfft_op    *d_ops;      /* device copy of s->ops */
} mfft_dsched;

int  mfft_dsched_upload(mfft_dsched *ds, mfft_sched *s);   /* takes ownership of s */
void mfft_dsched_free(mfft_dsched *ds);
/* run every stage of the schedule */
int  mfft_dsched_run(const mfft_dsched *ds, limb_t *slab, const mfft_geom *g,
                     const mfft_batch *d_batch, uint32_t nbatch, void *stream);

void *mfft_upload(const void *h, size_t bytes);            /* cudaMalloc + sync H2D; NULL on error */

/* One MFA transform (forward or inverse, truncated or not) as two passes + a finalize step.
 * Forward: mul_fft.c:2021-2068 / 2357-2409.  Inverse: 2411-2459 / 2925-2979. */
typedef struct {
   int inverse, truncated;
   uint64_t n, w, n1, n2, trunc_rows, N;      /* N = 2n blocks per slab half */
   uint32_t l, pitch, depth1, depth2;          /* depth1 = log2 n2, depth2 = log2 n1 */
   mfft_dsched col, row;
   mfft_batch *d_colb, *d_rowb; uint32_t ncolb, nrowb;
   mfft_move  *d_moves; uint32_t nmoves;
   uint32_t   *d_dst_base; uint32_t dst_stride;
   mfft_geom   gcol, grow;
   uint32_t   *rows;                           /* [nrows] logical valid rows, host */
   uint32_t    nrows;
   /* fused mode: the passes of each schedule, uploaded (tile.c / k_run_tiles) */
   int fused; uint32_t final_shift; int normalise;
   int fuse_split_ok;                          /* every input block is first read by the first column pass */
   mfft_passes pcol, prow;
   struct mfft_dpass { mfft_tile *d_tiles; uint32_t *d_pos; mfft_tileop *d_ops; uint32_t *d_stoff; } *dcol, *drow;
   uint32_t *d_dstpos; uint8_t *h_must_store; uint32_t *h_dstpos;
   /* sqrt2 transforms (FFT/IFFT_radix2_mfa_truncate_sqrt2): 4n coefficients, a column is 2*n2 positions
      (first-half rows, then second-half rows); for odd w the even and the odd columns have their own
      column schedules (class 0 = col/pcol/dcol/colb above, class 1 = the fields below) */
   int sqrt2, nclass; uint64_t trunc2;
   mfft_dsched col2; mfft_sched *h_col2; mfft_passes pcol2; struct mfft_dpass *dcol2;
   mfft_batch *d_colb2, *h_colb2; uint32_t ncolb2;
   uint32_t *d_dst_base2, *h_dst_base2; uint32_t ndst2;
   /* host copies of the tables (kept for the CPU-side schedule tests) */
   mfft_sched *h_col, *h_row;                  /* owned by col/row once uploaded */
   mfft_batch *h_colb, *h_rowb; mfft_move *h_moves; uint32_t *h_dst_base; uint32_t ndst;
} mfft_mfa;

/* trunc = 0: untruncated.  Returns 0 or a negative MPIRFFT_* code.
 * mfft_mfa_plan builds the schedules and tables on the host only (no device needed);
 * mfft_mfa_upload copies them to the device; mfft_mfa_build does both. */
/* final_shift: extra factor 2^final_shift (bit exponent mod 2NW) on every output; normalise:
 * outputs in canonical form.  mode: 0 = fused shared-memory passes when the coefficient size
 * allows (MPIRFFT_UNFUSED=1 in the environment overrides), 1 = one launch per radix-2 stage. */
int  mfft_mfa_plan(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                   uint32_t final_shift, int normalise, int mode);
/* nz (forward transforms): only the first nz input coefficients (reference order) can be non-zero -- the
 * split of an operand, mul_fft.c:3234-3236; 0 = no such promise.  Set before planning. */
void mfft_mfa_promise_zero_inputs(uint64_t nz);
int  mfft_mfa_upload(mfft_mfa *m);
int  mfft_mfa_build(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                    uint32_t final_shift, int normalise);
/* the sqrt2 variant: transform length 4n, trunc in (2n, 4n] a multiple of 2*n1 */
int  mfft_mfa_plan_sqrt2(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                         uint32_t final_shift, int normalise, int mode);
int  mfft_mfa_build_sqrt2(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                          uint32_t final_shift, int normalise);
void mfft_mfa_free(mfft_mfa *m);
/* slab: 2N blocks, input in half 0 in reference order (ii[k] = block k); dst: N blocks */
int  mfft_mfa_exec(const mfft_mfa *m, limb_t *slab, limb_t *dst, void *stream);
uint64_t mfft_mfa_launches(const mfft_mfa *m);
/* forward transform with FFT_split_bits fused into the first column pass (fused plans only) */
int  mfft_mfa_can_fuse_split(const mfft_mfa *m);
int  mfft_mfa_exec_split(const mfft_mfa *m, limb_t *slab, limb_t *dst, const limb_t *src, uint64_t nlimbs,
                         uint64_t bits, uint64_t ncoef, void *stream);

#ifdef __cplusplus
}
#endif
#endif
