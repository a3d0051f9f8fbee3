/* mpirfft_b200.h -- public C ABI of libmpirfft_b200.so
 *
 * A B200-native (sm_100a) implementation of the Schoenhage-Strassen multiplication path of
 * wbhart/mpir-fft, behind the reference's own C entry points.  The "plugin interface" of the
 * reference is simply the external linkage of mul_fft.c (every function is a global C symbol;
 * mul_fft.h:73-79 declares two of them).  Part 1 of this header re-declares those symbols with
 * byte-identical signatures, each citing the reference definition it replaces; linking a driver
 * against this library instead of mul_fft.o is the drop-in.  Part 2 is the device-resident
 * API the drop-in symbols are built on (operands and slabs stay in HBM between stages).
 *
 * Types: mp_limb_t = unsigned long (64-bit), mp_size_t = long, mp_bitcnt_t = unsigned long,
 * exactly as MPIR 2.4.0 defines them on LP64.  A coefficient block is (l+1) limbs: l limbs
 * of body and one signed two's-complement carry limb (README:54), value taken mod 2^(64 l)+1.
 *
 * Error behaviour: the reference has no error channel (void returns; illegal parameters
 * "will just segfault", mul_fft.c:3186-3187).  The drop-in symbols validate their parameters and
 * the device state and abort() with a diagnostic on stderr.  There is NO CPU fallback: without
 * a CUDA device every entry point that computes aborts (part 1) or returns an error (part 2).
 */
#ifndef MPIRFFT_B200_H
#define MPIRFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MPIRFFT_HAVE_MP_TYPES
typedef unsigned long mp_limb_t;
typedef long          mp_size_t;
typedef unsigned long mp_bitcnt_t;
#endif

/* ============================ Part 1: reference entry points ============================ */

/* mul_fft.c:3190  r1[0..n1+n2) = {i1,n1} * {i2,n2}; n = 2^depth, p = 2^(n w)+1.
 * (Implements the corrected row selection depth+1-depth/2 of mul_fft.c:3246.) */
void new_mpn_mul(mp_limb_t *r1, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2,
                 mp_bitcnt_t depth, mp_bitcnt_t w);

/* mul_fft.c:3573  the same product through the sqrt2 transforms: transform length 4n = 2^(depth+2)
 * with the 4n-th root of unity sqrt2^w, sqrt2 = 2^(3nw/4) - 2^(nw/4), so a ring of nw bits carries
 * twice as many coefficients as in new_mpn_mul.  Needs 2^(depth+1) < j1+j2-1 <= 2^(depth+2). */
void new_mpn_mul6(mp_limb_t *r1, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2,
                  mp_bitcnt_t depth, mp_bitcnt_t w);

/* mul_fft.c:41, 3119  r = i1*i2 mod 2^bits+1; c bit0 <=> i1 == 2^bits, bit1 <=> i2 == 2^bits;
 * returns 1 iff the result is 2^bits.  tt (2*(bits/64+1) limbs) is accepted and unused. */
mp_limb_t new_mpn_mulmod_2expp1(mp_limb_t *r, mp_limb_t *i1, mp_limb_t *i2, mp_limb_t c,
                                mp_limb_t bits, mp_limb_t *tt);

/* mul_fft.c:3125  r = i1*i2 mod 2^(n w)+1 on (n w/64 + 1)-limb blocks; returns the top bit
 * (the reference always returns 0 above 250 limbs, TODO:213-221; this returns the true bit). */
mp_limb_t fft_mulmod_2expp1(mp_limb_t *r, mp_limb_t *i1, mp_limb_t *i2, mp_size_t n, mp_size_t w,
                            mp_limb_t *tt);

/* mul_fft.c:2998  r1[0..r_limbs] = i1*i2 mod 2^(64 r_limbs)+1 for r_limbs-limb operands */
void FFT_mulmod_2expp1(mp_limb_t *r1, mp_limb_t *i1, mp_limb_t *i2, mp_size_t r_limbs,
                       mp_bitcnt_t depth, mp_bitcnt_t w);

/* mul_fft.c:2021 / 2411  length-2n MFA transform on n2 = 2n/n1 rows of n1 columns */
void FFT_radix2_mfa(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                    mp_limb_t **temp, mp_size_t n1);
void IFFT_radix2_mfa(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                     mp_limb_t **temp, mp_size_t n1);
/* mul_fft.c:2357 / 2925  truncated MFA (trunc a multiple of 2*n1) */
void FFT_radix2_mfa_truncate(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                             mp_limb_t **t2, mp_limb_t **temp, mp_size_t n1, mp_size_t trunc);
void IFFT_radix2_mfa_truncate(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                              mp_limb_t **t2, mp_limb_t **temp, mp_size_t n1, mp_size_t trunc);

/* mul_fft.c:2212 / 2593  length-4n truncated MFA with the sqrt2 trick: ii has 4n blocks (two halves of
 * n2 = 2n/n1 rows of n1 columns), trunc a multiple of 2*n1 in (2n, 4n]; temp is scratch in the
 * reference and unused here */
void FFT_radix2_mfa_truncate_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                                   mp_limb_t **t2, mp_limb_t **temp, mp_size_t n1, mp_size_t trunc);
void IFFT_radix2_mfa_truncate_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                                    mp_limb_t **t2, mp_limb_t **temp, mp_size_t n1, mp_size_t trunc);
/* mul_fft.c:2078 / 2461  the untruncated pair (= trunc 4n) */
void FFT_radix2_mfa_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                          mp_limb_t **temp, mp_size_t n1);
void IFFT_radix2_mfa_sqrt2(mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2,
                           mp_limb_t **temp, mp_size_t n1);

/* mul_fft.c:786 / 1444  length-2n radix-2 FFT (output bit-reversed) and its inverse (unscaled) */
void FFT_radix2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp);
void IFFT_radix2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                 mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp);
/* mul_fft.c:1128, 1028, 1674, 1538  truncated variants.  (rr, rs): result k is written into the block
 * rr[k*rs] points to (every caller of the reference passes rr == ii, rs == 1) */
void FFT_radix2_truncate(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                         mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc);
void FFT_radix2_truncate1(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                          mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc);
void IFFT_radix2_truncate(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                          mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc);
void IFFT_radix2_truncate1(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                           mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc);
/* mul_fft.c:839, 1488, 1230, 1792  length-4n transforms with the 4n-th root of unity sqrt2^w (trunc even,
 * in (2n, 4n]; even w falls back to the plain transforms with root 2^(w/2), as in the reference) */
void FFT_radix2_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                      mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp);
void IFFT_radix2_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                       mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp);
void FFT_radix2_truncate_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                               mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc);
void IFFT_radix2_truncate_sqrt2(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                                mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t trunc);
/* mul_fft.c:1397, 1964 (mul_fft.h:73-79), 1179, 1076, 1733, 1604  strided + twisted variants */
void FFT_radix2_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                        mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c,
                        mp_size_t rs);
void IFFT_radix2_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w, mp_limb_t **t1,
                         mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws, mp_size_t r, mp_size_t c,
                         mp_size_t rs);
void FFT_radix2_truncate_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w,
                                 mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws,
                                 mp_size_t r, mp_size_t c, mp_size_t rs, mp_size_t trunc);
void FFT_radix2_truncate1_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w,
                                  mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws,
                                  mp_size_t r, mp_size_t c, mp_size_t rs, mp_size_t trunc);
void IFFT_radix2_truncate_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w,
                                  mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws,
                                  mp_size_t r, mp_size_t c, mp_size_t rs, mp_size_t trunc);
void IFFT_radix2_truncate1_twiddle(mp_limb_t **ii, mp_size_t is, mp_size_t n, mp_bitcnt_t w,
                                   mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp, mp_size_t ws,
                                   mp_size_t r, mp_size_t c, mp_size_t rs, mp_size_t trunc);
/* mul_fft.c:1290, 1861  negacyclic transforms (even w: twist by powers of two; odd w: by powers of sqrt2^w) */
void FFT_radix2_negacyclic(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n, mp_bitcnt_t w,
                           mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp);
void IFFT_radix2_negacyclic(mp_limb_t **rr, mp_size_t rs, mp_limb_t **ii, mp_size_t n,
                            mp_bitcnt_t w, mp_limb_t **t1, mp_limb_t **t2, mp_limb_t **temp);

/* mul_fft.c:553, 639, 517, 721, 926  single butterflies / twiddle (outputs must not alias inputs) */
void FFT_radix2_butterfly(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t i,
                          mp_size_t n, mp_bitcnt_t w);
void FFT_radix2_inverse_butterfly(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2,
                                  mp_size_t i, mp_size_t n, mp_bitcnt_t w);
void FFT_radix2_twiddle_butterfly(mp_limb_t *u, mp_limb_t *v, mp_limb_t *s, mp_limb_t *t,
                                  mp_size_t NW, mp_bitcnt_t b1, mp_bitcnt_t b2);
void FFT_radix2_twiddle_inverse_butterfly(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2,
                                          mp_size_t NW, mp_bitcnt_t b1, mp_bitcnt_t b2);
void FFT_twiddle(mp_limb_t *r, mp_limb_t *i1, mp_size_t i, mp_size_t n, mp_bitcnt_t w);
/* mul_fft.c:591, 673, 972  the same with the 4n-th root of unity z1 = sqrt2^w (i and w odd) */
void FFT_radix2_butterfly_sqrt2(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2, mp_size_t i,
                                mp_size_t n, mp_bitcnt_t w, mp_limb_t *temp);
void FFT_radix2_inverse_butterfly_sqrt2(mp_limb_t *s, mp_limb_t *t, mp_limb_t *i1, mp_limb_t *i2,
                                        mp_size_t i, mp_size_t n, mp_bitcnt_t w, mp_limb_t *temp);
void FFT_twiddle_sqrt2(mp_limb_t *r, mp_limb_t *i1, mp_size_t i, mp_size_t n, mp_bitcnt_t w, mp_limb_t *temp);

/* mul_fft.c:272, 470, 494, 303, 394  arithmetic mod 2^(64 l)+1 on one block */
void mpn_normmod_2expp1(mp_limb_t *t, mp_size_t l);
void mpn_mul_2expmod_2expp1(mp_limb_t *t, mp_limb_t *i1, mp_size_t limbs, mp_bitcnt_t d);
void mpn_div_2expmod_2expp1(mp_limb_t *t, mp_limb_t *i1, mp_size_t limbs, mp_bitcnt_t d);
void mpn_lshB_sumdiffmod_2expp1(mp_limb_t *t, mp_limb_t *u, mp_limb_t *i1, mp_limb_t *i2,
                                mp_size_t limbs, mp_size_t x, mp_size_t y);
void mpn_sumdiff_rshBmod_2expp1(mp_limb_t *t, mp_limb_t *u, mp_limb_t *i1, mp_limb_t *i2,
                                mp_size_t limbs, mp_size_t x, mp_size_t y);

/* mul_fft.c:115, 87, 207, 180  integer <-> polynomial */
mp_size_t FFT_split_bits(mp_limb_t **poly, mp_limb_t *limbs, mp_size_t total_limbs, mp_size_t bits,
                         mp_size_t output_limbs);
mp_size_t FFT_split(mp_limb_t **poly, mp_limb_t *limbs, mp_size_t total_limbs,
                    mp_size_t coeff_limbs, mp_size_t output_limbs);
void FFT_combine_bits(mp_limb_t *res, mp_limb_t **poly, mp_size_t length, mp_size_t bits,
                      mp_size_t output_limbs, mp_size_t total_limbs);
void FFT_combine(mp_limb_t *res, mp_limb_t **poly, mp_size_t length, mp_size_t coeff_limbs,
                 mp_size_t output_limbs, mp_size_t total_limbs);

/* mul_fft.c:63  bit reversal (pure host arithmetic) */
mp_limb_t mpir_revbin(mp_limb_t in, mp_bitcnt_t bits);

/* ============================ Part 2: device-resident API ============================== */

#define MPIRFFT_OK          0
#define MPIRFFT_ENODEV     (-1)   /* no usable CUDA device / CUDA error (see mpirfft_last_error) */
#define MPIRFFT_EINVAL     (-2)   /* illegal parameters (A.5 of SURVEY.md) */
#define MPIRFFT_ENOMEM     (-3)

/* Select the device for this process (default 0) and create the context.  Idempotent. */
int  mpirfft_init(int device);
int  mpirfft_device_count(void);
const char *mpirfft_last_error(void);
const char *mpirfft_version(void);

/* Parameter legality of new_mpn_mul (mul_fft.c:3193-3203); fills the derived sizes. */
typedef struct {
   uint64_t n, bits1, sqrt, j1, j2, trunc, limbs, n2, trunc_rows;
} mpirfft_mul_params;
int  mpirfft_mul_params_get(mpirfft_mul_params *out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth,
                            mp_bitcnt_t w);
/* the same for new_mpn_mul6 (mul_fft.c:3576-3612): n2 = rows per half, trunc_rows = live rows of both halves */
int  mpirfft_mul6_params_get(mpirfft_mul_params *out, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth,
                             mp_bitcnt_t w);
/* Parameter chooser (the reference has none, mul_fft.c:3177-3178): the smallest coefficient ring
 * among 64..512 limbs that is legal for an n1 x n2 limb product (the fused tile executor and the
 * warp-level product kernel), else the smallest legal (depth, w) with w in {1,2}. */
int  mpirfft_choose_params(mp_size_t n1, mp_size_t n2, mp_bitcnt_t *depth, mp_bitcnt_t *w);
/* the same with the new_mpn_mul6 shape as a candidate at every ring size (*sqrt2 = 1: use new_mpn_mul6 /
 * mpirfft_mul6_plan_create): what mpirfft_mpn_mul uses */
int  mpirfft_choose_params6(mp_size_t n1, mp_size_t n2, mp_bitcnt_t *depth, mp_bitcnt_t *w, int *sqrt2);
/* r[0..n1+n2) = i1 * i2 with the parameters of mpirfft_choose_params: the mpn_mul-shaped entry the
 * reference leaves as a FIXME (mul_fft.c:3177-3178).  Host pointers; aborts like new_mpn_mul. */
void mpirfft_mpn_mul(mp_limb_t *r, mp_limb_t *i1, mp_size_t n1, mp_limb_t *i2, mp_size_t n2);

/* A multiplication plan owns the schedules and the HBM slabs for one (n1, n2, depth, w).  With a
 * coefficient ring above 512 limbs (2^depth * w > 32768) the plan is the plan of the sharded
 * multiplication on one rank (mpirfft_smul_*, below): it executes as a whole -- mpirfft_mul_exec_phase
 * returns MPIRFFT_EINVAL, the byte and launch counters read 0. */
typedef struct mpirfft_mul_plan mpirfft_mul_plan;
int  mpirfft_mul_plan_create(mpirfft_mul_plan **plan, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth,
                             mp_bitcnt_t w);
/* the plan of new_mpn_mul6 (sqrt2 transforms); executed and destroyed like any other plan */
int  mpirfft_mul6_plan_create(mpirfft_mul_plan **plan, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth,
                              mp_bitcnt_t w);
void mpirfft_mul_plan_destroy(mpirfft_mul_plan *plan);
/* d_r[0..n1+n2) = d_i1 * d_i2, all three device pointers; asynchronous on `stream`
 * (a cudaStream_t, NULL = default stream). */
int  mpirfft_mul_exec_device(mpirfft_mul_plan *plan, mp_limb_t *d_r, const mp_limb_t *d_i1,
                             const mp_limb_t *d_i2, void *stream);
/* same with host operands: H2D copies, the product, D2H copy, synchronised on return */
int  mpirfft_mul_exec_host(mpirfft_mul_plan *plan, mp_limb_t *r, const mp_limb_t *i1,
                           const mp_limb_t *i2);
/* run only one stage group of the pipeline on the plan's internal buffers (bench / profiling):
 * 0 split+fft(i1), 1 split+fft(i2), 2 pointwise, 3 ifft, 4 combine */
int  mpirfft_mul_exec_phase(mpirfft_mul_plan *plan, int phase, mp_limb_t *d_r,
                            const mp_limb_t *d_i1, const mp_limb_t *d_i2, void *stream);
size_t mpirfft_mul_plan_device_bytes(const mpirfft_mul_plan *plan);
/* number of kernel launches one product issues */
uint64_t mpirfft_mul_plan_launches(const mpirfft_mul_plan *plan);

/* Batched pointwise products on device: for k < count,
 *   d_a[k*pitch .. +l] = d_a[k] * d_b[k] mod 2^(64 l)+1   (blocks canonical: top 0, or top 1 & body 0) */
int  mpirfft_mulmod_batch_device(mp_limb_t *d_a, const mp_limb_t *d_b, size_t count, size_t l,
                                 size_t pitch, void *stream);

/* ---- batched mulmod 2^(64 L)+1 through a negacyclic transform (FFT_mulmod_2expp1, mul_fft.c:2998;
 * parameter choice after fft_mulmod_2expp1, 3125-3167).  `count` independent products per call,
 * operands and results resident in HBM: blocks of L+1 limbs (canonical: top limb 0, or 1 with
 * zero limbs) at `pitch` limbs; r may alias a.  Shapes: 2^depth * w = 128 * L / 2^(depth+1), w
 * even, L / 2^(depth+1) a multiple of 32 (mpirfft_mulmod_params picks one for L). */
typedef struct mpirfft_mulmod_plan mpirfft_mulmod_plan;
int  mpirfft_mulmod_params(mp_size_t r_limbs, mp_bitcnt_t *depth, mp_bitcnt_t *w);
int  mpirfft_mulmod_plan_create(mpirfft_mulmod_plan **plan, mp_size_t r_limbs, mp_bitcnt_t depth, mp_bitcnt_t w, size_t count);
void mpirfft_mulmod_plan_destroy(mpirfft_mulmod_plan *plan);
size_t mpirfft_mulmod_plan_device_bytes(const mpirfft_mulmod_plan *plan);
/* phase < 0: the whole product; 0..4: split+forward(a), split+forward(b), pointwise, inverse, finish */
int  mpirfft_mulmod_plan_exec(mpirfft_mulmod_plan *plan, mp_limb_t *d_r, const mp_limb_t *d_a, const mp_limb_t *d_b,
                              size_t pitch, int phase, void *stream);

/* ---- sharded multiplication: the MFA spread over the GPUs of one box (one process per GPU) ----
 * Columns are block-partitioned for the column passes, the live rows for the row passes and the
 * pointwise products.  The library runs the local phases on plan-owned HBM buffers; the host
 * language issues the collectives on the exchange buffers named in the layout (counts in limbs):
 *   phase 0 (which, d_in)  split + column FFTs of operand `which` -> send
 *        A: all_to_all send -> recv;  to rank h: rows_h*ncl blocks at offset r0_h*ncl blocks
 *   phase 1 (which)        unpack recv, row FFTs -> spectrum of operand `which`
 *   phase 2                pointwise products
 *   phase 3                row IFFTs -> send;  B: all_to_all send -> work; to rank g: nrl*ncl blocks at g*nrl*ncl
 *   phase 4                column IFFTs, scaled and normalised -> send;  C: same as A
 *   phase 5                unpack recv behind the halo slots; H: every rank > 0 copies the last
 *                          halo_send blocks of rank-1 (mpirfft_smul_tail_blocks) into unp[0..)
 *   phase 6                recombine my window of the result (limbs [limb_lo, limb_hi)) -> out
 *   mpirfft_smul_carry     rank 0 .. world-1 in turn: add the carry handed over, hand on the new one */
typedef struct mpirfft_smul_plan mpirfft_smul_plan;
typedef struct {
   uint32_t block_limbs, ncl, nrl, r0, trunc_rows, n1cols, halo, halo_send;
   uint64_t limb_lo, limb_hi, send_limbs, recv_limbs, work_limbs;
   void *send, *recv, *work, *unp, *out;
} mpirfft_smul_layout;
int  mpirfft_smul_plan_create(mpirfft_smul_plan **plan, mp_size_t n1, mp_size_t n2, mp_bitcnt_t depth,
                              mp_bitcnt_t w, int rank, int world);
void mpirfft_smul_plan_destroy(mpirfft_smul_plan *plan);
int  mpirfft_smul_info(const mpirfft_smul_plan *plan, mpirfft_smul_layout *out);
int  mpirfft_smul_phase(mpirfft_smul_plan *plan, int phase, int which, const mp_limb_t *d_in, void *stream);
int  mpirfft_smul_carry(mpirfft_smul_plan *plan, unsigned carry_in, unsigned *carry_out, void *stream);
mp_limb_t *mpirfft_smul_tail_blocks(mpirfft_smul_plan *plan, unsigned count);

/* Pointwise algorithm (new_mpn_mulmod_2expp1, mul_fft.c:3119): 0 = default (IMAD.WIDE carry-chain
 * kernel; its block products are Karatsuba-split at l = 128 and 256, schoolbook otherwise),
 * 1 = nested Schoenhage-Strassen step inside a warp (the fft_mulmod_2expp1 idea, mul_fft.c:3125-3167)
 * for l in {64,128,256,512}, 2 = Karatsuba-split blocks wherever built (l = 64, 128, 256),
 * 3 = schoolbook blocks everywhere.  Every mode is bit-exact; the choice only moves time.
 * Environment: MPIRFFT_POINTWISE = s | k | d. */
void mpirfft_set_pointwise_mode(int mode);

/* device memory helpers (so that a host language needs nothing but this ABI) */
void *mpirfft_malloc_device(size_t bytes);
void  mpirfft_free_device(void *p);
void *mpirfft_malloc_pinned(size_t bytes);
void  mpirfft_free_pinned(void *p);
int   mpirfft_memcpy_h2d(void *d, const void *h, size_t bytes, void *stream);
int   mpirfft_memcpy_d2h(void *h, const void *d, size_t bytes, void *stream);
int   mpirfft_stream_sync(void *stream);

/* Optional per-kernel-class timing with CUDA events on the launching stream (used by bench.py
 * for the roofline figures; off by default).  Classes: 0 transform stage, 1 finalize/normalise
 * gather, 2 pointwise, 3 split, 4 combine, 5 normalise.  read() drains and sums the records:
 * total ms, launches and algorithmic bytes (stage class only) per class. */
void mpirfft_profile_enable(int on);
int  mpirfft_profile_read(double *ms, uint64_t *launches, double *bytes, int nclass);

/* Integer multiply-accumulate issue rate of this GPU in MAD/s (the denominator of the pointwise
 * kernel's roofline): mode 0 IMAD.WIDE.U32, 1 IMAD.WIDE.U32.X carry chains, 2 32-bit IMAD. */
double mpirfft_measure_imad_rate(int mode);

/* launches issued through the library since the last reset (bench.py's gpu_launches) */
uint64_t mpirfft_launch_count(void);
void     mpirfft_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif
