/* sched.h -- host-side (plain C) schedule generator: turns the reference's recursive radix-2
 * transforms into data-independent op lists (see mfft_internal.h for the op format). */
#ifndef MFFT_SCHED_H
#define MFFT_SCHED_H

#include "../mfft_internal.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
   uint32_t  S;          /* logical positions */
   uint64_t  NW;         /* ring is 2^NW + 1 */
   uint64_t  M2;         /* 2*NW: exponents live mod M2 */
   uint32_t *slot;       /* [S]  current slot of each logical position (in [0,2S)) */
   uint32_t *wr_stage;   /* [2S] last stage that wrote the slot */
   uint32_t *rd_stage;   /* [2S] last stage that read the slot */
   mfft_op  *ops;        /* emitted ops; after mfft_sched_finish sorted by stage */
   size_t    nops, cap;
   uint32_t  nstages;    /* valid after finish; stages are numbered 1..nstages */
   uint32_t *stage_off;  /* [nstages+1] offsets into ops (after finish) */
   /* position-based view (fused executor): phys[k] = physical position holding logical k */
   uint32_t *phys;       /* [S] */
   uint32_t *pwr_stage;  /* [S] last pstage that wrote the physical position */
   uint32_t *prd_stage;  /* [S] last pstage that read it */
   uint32_t  npstages;
   /* known-zero tracking (optional): zero[k] != 0 = logical position k currently holds 0.  emit() then
      drops zero operands, turns "x + 0 in place" into nothing and ops on zeros into zeros */
   uint8_t  *zero;
} mfft_sched;

typedef enum {
   MFFT_T_FFT = 0,          /* FFT_radix2 / FFT_radix2_twiddle            mul_fft.c:786, 1397  */
   MFFT_T_FFT_TRUNC,        /* FFT_radix2_truncate(_twiddle)             mul_fft.c:1128, 1179 */
   MFFT_T_FFT_TRUNC1,       /* FFT_radix2_truncate1(_twiddle)            mul_fft.c:1028, 1076 */
   MFFT_T_IFFT,             /* IFFT_radix2 / IFFT_radix2_twiddle          mul_fft.c:1444, 1964 */
   MFFT_T_IFFT_TRUNC,       /* IFFT_radix2_truncate(_twiddle)            mul_fft.c:1674, 1733 */
   MFFT_T_IFFT_TRUNC1,      /* IFFT_radix2_truncate1(_twiddle)           mul_fft.c:1538, 1604 */
   MFFT_T_FFT_NEGACYCLIC,   /* FFT_radix2_negacyclic (even or odd w)     mul_fft.c:1290 */
   MFFT_T_IFFT_NEGACYCLIC   /* IFFT_radix2_negacyclic (even or odd w)    mul_fft.c:1861 */
} mfft_transform_kind;

/* S = number of logical positions, NW = n*w of the ring */
mfft_sched *mfft_sched_new(uint32_t S, uint64_t NW);
void        mfft_sched_free(mfft_sched *s);

/* Emit the ops of one length-2n transform over positions p0, p0+is, ..., root of unity 2^w,
 * per-column twist 2^{ws*c*(r + rs*f)} on frequency f (ws = 0: untwisted), truncation `trunc`
 * (ignored by the untruncated kinds).  Returns 0, or <0 on illegal parameters. */
int  mfft_sched_emit(mfft_sched *s, mfft_transform_kind kind, uint32_t p0, uint32_t is,
                     uint64_t n, uint64_t w, uint64_t ws, uint64_t r, uint64_t rs, uint64_t trunc);

/* The column schedule of the sqrt2 MFA (FFT/IFFT_radix2_mfa_truncate_sqrt2, mul_fft.c:2212-2355,
 * 2593-2750) over the 2*n2 positions of one column (first-half rows, then second-half rows, stride 1):
 * the layer between the halves with the sqrt2^w twiddles, the two twisted column transforms (the
 * second one truncated to trunc2 rows) and the row relabels.  par: 0 / 1 = the op list of the even /
 * odd columns (w odd; batch entries carry col with column = 2 col + par), -1 = all columns (w even).
 * pad: the even columns get a copy op where the odd ones have their extra op (ping-pong view: both
 * classes must leave every position in the same slab half). */
int  mfft_sched_emit_sqrt2_cols(mfft_sched *s, int inverse, uint64_t n2, uint64_t n1, uint64_t w, uint64_t trunc2, int par, int pad);

/* the 1-D transforms of length 4n with the root sqrt2^w (mul_fft.c:839, 1230, 1488, 1792); S = 4n */
int  mfft_sched_emit_sqrt2_1d(mfft_sched *s, int inverse, uint64_t n, uint64_t w, uint64_t trunc);

/* Emit one explicit op: position pS <- sSA*A*2^eSA + sSB*B*2^eSB and (optionally) position
 * pT <- sTA*A*2^eTA + sTB*B*2^eTB, A/B = current contents of positions posA/posB (posB and pT
 * may be MFFT_NONE).  Exponents are bit counts mod 2*NW. */
void mfft_sched_emit_op(mfft_sched *s, uint32_t posA, uint32_t posB,
                        uint32_t pS, int sSA, uint64_t eSA, int sSB, uint64_t eSB,
                        uint32_t pT, int sTA, uint64_t eTA, int sTB, uint64_t eTB);

/* Declare the logical positions first, first+1, ... as known zeros before emitting a transform: an
 * operand of new_mpn_mul fills only about half of the truncated length (j1 of trunc coefficients,
 * mul_fft.c:3234-3236 zero-fills the rest), so the first layers of the forward transform are mostly
 * "x + 0" and "(x - 0) 2^e". */
int  mfft_sched_zero_from(mfft_sched *s, uint32_t first);

/* swap the slots of logical positions a and b (the reference's pointer swaps, e.g. 2380-2389) */
void mfft_sched_swap(mfft_sched *s, uint32_t a, uint32_t b);

/* bit-reverse relabel of positions p0 + is*j, j < 2^bits (mul_fft.c:2041-2050 and twins) */
void mfft_sched_revbin(mfft_sched *s, uint32_t p0, uint32_t is, uint32_t bits);

/* sort ops by stage and build stage_off */
int  mfft_sched_finish(mfft_sched *s);

uint64_t mfft_revbin(uint64_t in, uint32_t bits);   /* mpir_revbin, mul_fft.c:63-79 */

#ifdef __cplusplus
}
#endif
#endif
