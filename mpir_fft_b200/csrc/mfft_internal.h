/* mfft_internal.h -- shared between the C host side (csrc/host) and the thin CUDA launch
 * ABI (csrc/cuda).  Not a public header; the public C ABI is include/mpirfft_b200.h.
 *
 * Vocabulary (follows the reference, /root/reference/mul_fft.c:44-50 and README:48-60):
 *   p = 2^NW + 1, NW = n*w bits, l = NW/64 limbs; a *coefficient block* is l limbs of body plus
 *   one signed carry limb ("top"), value = body + top*2^NW  (== body - top mod p).
 *   Coefficient blocks live in one contiguous HBM *slab* with a fixed pitch (in limbs).
 *
 * The reference permutes host pointers (README:56); here every logical position has two
 * candidate *slots* (slab half 0 and half 1, "ping-pong") and a transform is a data-independent
 * list of *ops* grouped into *stages*; ops of one stage are independent and run as one launch.
 */
#ifndef MFFT_INTERNAL_H
#define MFFT_INTERNAL_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t limb_t;

/* pitch (limbs) of the coefficient blocks in library-owned HBM slabs: l body limbs + the top limb,
 * rounded up to an even count so that every block starts 16-byte aligned (128-bit global accesses,
 * 16-byte asynchronous copies into shared memory) */
static inline uint32_t mfft_pitch(uint32_t l) { return (l + 2u) & ~1u; }

#define MFFT_NONE 0xFFFFFFFFu

/* One op: up to two outputs, each a signed sum of up to two inputs multiplied by powers of two:
 *   outS = sSA * A * 2^(eSA + col*cSA) + sSB * B * 2^(eSB + col*cSB)      (mod p)
 *   outT = sTA * A * 2^(eTA + col*cTA) + sTB * B * 2^(eTB + col*cTB)      (mod p)
 * All exponents are bit counts reduced mod 2*NW (2^NW == -1).  `col` is the per-batch-entry
 * column index (the MFA twist z^{r*c}, mul_fft.c:1409-1411); it is 0 for untwisted passes.
 * inA/inB/outS/outT are slots in [0, 2*S).  sXY == 0 means "term absent". */
typedef struct {
   uint32_t inA, inB, outS, outT;
   uint32_t eSA, eSB, eTA, eTB;
   uint32_t cSA, cSB, cTA, cTB;
   int8_t   sSA, sSB, sTA, sTB;
   uint32_t stage;
   /* the same op in terms of PHYSICAL positions (blocks of slab half 0 that never move; the
    * relabelling of the reference's pointer swaps is folded in) and its stage under in-place
    * hazards -- used by the fused shared-memory tile executor (tile.c, k_run_tiles) */
   uint32_t pA, pB, pS, pT;
   uint32_t pstage;
   /* shape of the op for the carry-save stage kernel (xform.c: classify; MFFT_K_* below) */
   uint32_t kind, kparam;
} mfft_op;

/* One batch entry: which strided family of blocks an op list is applied to. */
typedef struct {
   uint32_t base;    /* block index of position 0 (before slot stride) inside a slab half */
   uint32_t parity;  /* 0/1: xor'ed into the slab-half bit of every slot */
   uint32_t col;     /* column index used for the twist exponents */
   uint32_t pad;
} mfft_batch;

/* Geometry of one pass over a slab. block index of (slot, batch b) =
 *   (((slot / S) ^ parity_b) * half_blocks) + base_b + (slot % S) * slot_stride            */
typedef struct {
   uint32_t S;            /* positions in the pass */
   uint32_t slot_stride;  /* in blocks */
   uint64_t half_blocks;  /* blocks per slab half */
   uint32_t l;            /* limbs of body */
   uint32_t pitch;        /* limbs per block in the slab (>= l+1) */
} mfft_geom;

/* ---- fused tile passes: several stages executed inside shared memory ---------------------- */
#define MFFT_TILE_LOAD  0x80000000u      /* flags in the position list */
#define MFFT_TILE_STORE 0x40000000u
#define MFFT_TILE_POSMASK 0x3FFFFFFFu

typedef struct {            /* one op of a tile; a,b,s,t index the tile's position list */
   uint16_t a, b, s, t;     /* 0xFFFF = none */
   uint32_t eSA, eSB, eTA, eTB, cSA, cSB, cTA, cTB;
   int8_t   sSA, sSB, sTA, sTB;
   uint32_t lstage;         /* stage inside the pass, 0-based, ops sorted by it */
   /* host-side classification of untwisted ops whose rotations are all multiples of 128 bits
      (tile.c: classify_op), so that the kernel runs the common butterfly shapes without decoding
      exponents: kparam = chunk rotation yc (e / 128 after folding e >= NW into the sign) | neg << 31 */
   uint32_t kind;           /* MFFT_K_* */
   uint32_t kparam;
   uint32_t pad[2];         /* 64 bytes: descriptors are fetched with 16-byte asynchronous copies */
} mfft_tileop;

#define MFFT_K_ANY   0u     /* decode at run time (twisted or unaligned ops, rare shapes) */
#define MFFT_K_FWD   1u     /* S = A + B,             T = +-(A - B) * 2^(128 yc)      (553-576)  */
#define MFFT_K_INV   2u     /* S = A +- B * 2^(128 yc), T = A -+ B * 2^(128 yc)       (639-660)  */
#define MFFT_K_ROT   3u     /* S = +-A * 2^(128 yc)                                   (926-957)  */
#define MFFT_K_ADD   4u     /* S = A + B                                              (1093)     */
/* two consecutive layers fused on four positions (a, b, s, t = P0..P3; eSA, eSB, eTA, eTB = the four
   packed rotations y | neg << 31; tile.c: fuse_radix4, mfft_tiles.h: fwd4_unit / inv4_unit) */
#define MFFT_K_FWD4  5u
#define MFFT_K_INV4  6u
#define MFFT_K_NOP   7u     /* host only: op absorbed into a fused unit, removed before upload */
#define MFFT_K_DBL   8u     /* S = 2 A           (the doubling steps of IFFT_radix2_truncate, 1788-1789) */
#define MFFT_K_HALF  9u     /* S = (A + B) / 2   (the averaging steps of IFFT_radix2_truncate1, 1556-1560) */
#define MFFT_K_SHR   10u    /* S = A / 2^s, 1 <= s <= 31 = kparam   (the final 2^-(depth+1) scaling, 3256-3258) */
#define MFFT_K_2AMB  11u    /* S = 2A - B [, T = +-(A - B) * 2^(128 yc)]   (1630-1631, 1639-1647) */

/* pad = offset of the tile's (nstages+1) stage offsets in the pass's stoff array (<= 63 stages) */
typedef struct { uint32_t pos_off, npos, op_off, nops, nstages, pad; } mfft_tile;

/* gather/normalise descriptor (finalize step of a transform): for logical position k:
 * src slot (with batch parity/base as above) -> dst block index  dst_base_b + dst[k]*dst_stride */
typedef struct {
   uint32_t src_slot;
   uint32_t dst_pos;
} mfft_move;

/* ------------------------------------------------------------------------------------------
 * CUDA launch ABI (implemented in csrc/cuda/mfft_kernels.cu).  All pointers are device
 * pointers unless named h_*.  `stream` is a cudaStream_t passed as void*.  Every function
 * returns 0 on success or a negative mfft error code; none of them ever falls back to the CPU.
 * ------------------------------------------------------------------------------------------ */
int  mfft_dev_init(int device);                       /* select device, create context; <0 if no GPU */
int  mfft_dev_bind(void);                             /* make the device of mfft_dev_init current in the calling thread */
int  mfft_dev_count(void);
void *mfft_dev_alloc(size_t bytes);                   /* NULL on failure */
void mfft_dev_free(void *p);
void *mfft_host_alloc_pinned(size_t bytes);
void mfft_host_free_pinned(void *p);
int  mfft_dev_h2d(void *d, const void *h, size_t bytes, void *stream);
int  mfft_dev_d2h(void *h, const void *d, size_t bytes, void *stream);
int  mfft_dev_h2d_2d(void *d, size_t dpitch, const void *h, size_t hpitch, size_t width, size_t rows, void *stream);
int  mfft_dev_d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width, size_t rows, void *stream);
int  mfft_dev_d2d(void *d, const void *s, size_t bytes, void *stream);
int  mfft_dev_memset0(void *d, size_t bytes, void *stream);
int  mfft_dev_sync(void *stream);
void *mfft_dev_stream_create(void);                   /* non-blocking stream; NULL on failure */
void mfft_dev_stream_destroy(void *s);
void *mfft_dev_event_create(void);
void mfft_dev_event_destroy(void *e);
int  mfft_dev_event_record(void *e, void *stream);
int  mfft_dev_stream_wait(void *stream, void *e);      /* stream waits for the event */
int  mfft_dev_event_sync(void *e);                     /* the calling host thread waits for the event */
int  mfft_dev_host_is_pinned(const void *p);           /* 1: page-locked / registered (or device) memory, 0: pageable */
const char *mfft_dev_last_error(void);

/* run ops[first .. first+count) (one stage) over nbatch batch entries */
int  mfft_dev_run_stage(limb_t *slab, const mfft_geom *g, const mfft_op *d_ops, uint32_t count,
                        const mfft_batch *d_batch, uint32_t nbatch, void *stream);

/* Carry-save stage path for coefficient rings of any even size (mfft_cs_stage.h): the same chunk
 * arithmetic as the tile executor on HBM-resident blocks, one launch per layer; cw holds l/2 carry
 * words per block of the slab (both halves).  init: carry words of the S positions of half 0 of
 * every batch entry (zero, last = the block's top limb); run_stage: ops must carry a kind
 * (MFFT_K_* or MFFT_K_ANY with a single operand); finalize: gather + resolve (+ normalise). */
int  mfft_dev_cs_init(const limb_t *slab, int32_t *cw, const mfft_geom *g, const mfft_batch *d_batch, uint32_t nbatch, void *stream);
int  mfft_dev_run_stage_cs(limb_t *slab, int32_t *cw, const mfft_geom *g, const mfft_op *d_ops, uint32_t count,
                           const mfft_batch *d_batch, uint32_t nbatch, void *stream);
int  mfft_dev_finalize_cs(limb_t *dst, uint32_t dst_stride, const uint32_t *d_dst_base, limb_t *slab, int32_t *cw,
                          const mfft_geom *g, const mfft_move *d_moves, uint32_t nmoves, const mfft_batch *d_batch,
                          uint32_t nbatch, int normalise, void *stream);

/* split != NULL in mfft_dev_run_tiles: the launch cuts the coefficients it loads out of the operand
 * {src, nlimbs} (block k of slab half 0 = bits [k*bits, (k+1)*bits), zero for k >= ncoef;
 * FFT_split_bits, mul_fft.c:115-170 and the zero fill 3235-3236) instead of reading them from the slab */
typedef struct { const limb_t *src; uint64_t nlimbs, bits, ncoef; } mfft_split;

/* Big rings, in place on slab half 0 (position view).  run_stage_cs_ip: the ops of one stage, operands
 * staged in shared memory (ops carry pA/pB/pS/pT and a kind).  run_tiles_sliced: several layers on the
 * chunk slices { i0 + gs t } of the tile's coefficients (see mfft_cs_stage.h); the pass descriptors
 * are those of the tile executor for the virtual ring of l/2/gs chunks; a CTA takes R adjacent slices. */
int  mfft_dev_run_stage_cs_ip(limb_t *slab, int32_t *cw, const mfft_geom *g, const mfft_op *d_ops, uint32_t count,
                              const mfft_batch *d_batch, uint32_t nbatch, uint32_t nstaged, void *stream);
int  mfft_dev_run_tiles_sliced(limb_t *slab, int32_t *cw, const mfft_geom *g, uint32_t gs, uint32_t nchv, uint32_t R,
                               const mfft_tile *d_tiles, uint32_t ntiles, const uint32_t *d_pos, const mfft_tileop *d_ops,
                               const uint32_t *d_stoff, uint32_t max_npos, uint32_t max_nops,
                               const mfft_batch *d_batch, uint32_t nbatch, void *stream);
/* shared-memory bytes one coefficient of nchv chunks takes in a sliced tile; 0 if nchv is not built */
size_t mfft_dev_sliced_coeff_bytes(uint32_t nchv);

/* One fused pass: CTA (tile, batch entry) loads the tile's positions into shared memory, runs
 * all its stages there and stores the written positions back in place -- or, if dst != NULL,
 * to dst[(dst_base[b] + dstpos[i]*dst_stride)*pitch] (dstpos parallel to the position list,
 * MFFT_NONE = do not store), normalised when `normalise` is set.  max_npos bounds the tile size.
 * heavy: many ops of the pass need the run-time decoded general path (more registers per thread).
 * h_tiles / h_pos / h_stoff (optional): host copies of the tile descriptors; small passes get them
 * as kernel parameters, which takes two dependent global loads out of every CTA's start-up. */
int  mfft_dev_run_tiles(limb_t *slab, const mfft_geom *g, const mfft_tile *d_tiles, uint32_t ntiles,
                        const uint32_t *d_pos, const mfft_tileop *d_ops, uint32_t max_npos, uint32_t max_nops,
                        const mfft_batch *d_batch, uint32_t nbatch,
                        limb_t *dst, const uint32_t *d_dstpos, const uint32_t *d_dst_base,
                        uint32_t dst_stride, int normalise, const uint32_t *d_stoff, int heavy,
                        const mfft_tile *h_tiles, const uint32_t *h_pos, const uint32_t *h_stoff,
                        const mfft_batch *h_batch, const mfft_split *split, void *stream);
/* 1 if the fused executor supports coefficient size l (else use mfft_dev_run_stage) */
int  mfft_dev_tiles_supported(uint32_t l);
uint32_t mfft_dev_tiles_max_npos(uint32_t l);

/* dst[dst_base_b + mv.dst_pos*dst_stride] = normalise( src(slot mv.src_slot, batch b) * 2^shift )
 * shift is a bit exponent mod 2*NW (0 = none); if !normalise the block is copied unreduced. */
int  mfft_dev_finalize(limb_t *dst, uint32_t dst_stride, const uint32_t *d_dst_base,
                       const limb_t *slab, const mfft_geom *g, const mfft_move *d_moves, uint32_t nmoves,
                       const mfft_batch *d_batch, uint32_t nbatch, uint32_t shift, int normalise,
                       void *stream);

/* pointwise a[i] = a[i]*b[i] mod p over nblk canonical blocks listed in d_blocks (block indices) */
int  mfft_dev_pointwise(limb_t *a, const limb_t *b, const uint32_t *d_blocks, uint32_t nblk,
                        uint32_t l, uint32_t pitch, void *stream);

/* 0: schoolbook kernel (default), 1: nested Schoenhage-Strassen kernel where available */
void mfft_dev_pointwise_mode(int mode);

/* split: coefficient i (i < ncoef) = bits [i*bits, (i+1)*bits) of {src, nlimbs}, zero-extended to a
 * block (FFT_split_bits, mul_fft.c:115-170); blocks ncoef..nzero-1 are zeroed (mul_fft.c:3235). */
int  mfft_dev_split(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *src, uint64_t nlimbs,
                    uint64_t bits, uint64_t ncoef, uint64_t nzero, void *stream);

/* the same for a rank's column shard: local block (r, cl) = coefficient r*n1 + c0 + cl, r < nrows */
int  mfft_dev_split_cols(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *src, uint64_t nlimbs,
                         uint64_t bits, uint64_t ncoef, uint64_t nrows, uint32_t ncl, uint32_t n1, uint32_t c0, void *stream);
/* dst block i = src block d_table[i] */
int  mfft_dev_gather_blocks(limb_t *dst, const limb_t *src, const uint32_t *d_table, uint64_t nblocks, uint32_t pitch, void *stream);
/* res += c (small) with full ripple; *d_carry_out = carry leaving the array */
int  mfft_dev_add_small(limb_t *res, uint64_t total, uint32_t c, uint32_t *d_carry_out, void *stream);
/* windowed combine: res[k] = limb k of the window starting at bit base_bit of the full sum (carries
 * inside the window only); *d_carry_out (optional, device) = carry leaving limb total-1 */
int  mfft_dev_combine_window(limb_t *res, uint64_t total, const limb_t *slab, uint32_t l, uint32_t pitch,
                             uint64_t bits, uint64_t ncoef, uint64_t base_bit, uint32_t *d_carry_out, void *work, void *stream);

/* combine: res[0..total) = sum_{i<ncoef} block_i * 2^(i*bits) truncated to total limbs
 * (FFT_combine_bits, mul_fft.c:207-267, including the MPN_ZERO of mul_fft.c:3261).
 * blocks must be normalised with zero top limb.  work: >= mfft_dev_combine_work(total) bytes. */
size_t mfft_dev_combine_work(uint64_t total);
int  mfft_dev_combine(limb_t *res, uint64_t total, const limb_t *slab, uint32_t l, uint32_t pitch,
                      uint64_t bits, uint64_t ncoef, void *work, void *stream);

/* batched mulmod through a negacyclic transform (FFT_mulmod_2expp1, mul_fft.c:2998-3117):
 * split: block (prod, k), k < K, = limbs [k*pl, (k+1)*pl) of operand prod, zero-extended to l+1 limbs;
 * finish: r[prod] = sum_k c_k B^(pl k) mod B^(K pl) + 1, canonical, from the K canonical coefficients
 * C[prod*K + k] mod B^l + 1 (l = 2 pl) and the operands' low limbs (which fix c_k mod 2^64); operands
 * with a set top limb (== -1) are handled exactly.  r may alias a. */
int  mfft_dev_mm_split(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *src, uint32_t src_pitch,
                       uint32_t K, uint32_t pl, uint64_t count, void *stream);
int  mfft_dev_mm_finish(limb_t *r, uint32_t r_pitch, const limb_t *a, const limb_t *b, uint32_t ab_pitch,
                        const limb_t *C, uint32_t l, uint32_t c_pitch, uint32_t K, uint32_t pl, uint64_t count, void *stream);

/* normalise nblk blocks in place (mpn_normmod_2expp1, mul_fft.c:272-294) */
int  mfft_dev_normalise(limb_t *slab, uint32_t l, uint32_t pitch, uint64_t nblk, void *stream);

/* optional per-kernel-class CUDA-event timing (bench.py's roofline leg); classes in order:
 * 0 stage, 1 finalize, 2 pointwise, 3 split, 4 combine, 5 normalise */
void mfft_dev_profile_enable(int on);
int  mfft_dev_profile_is_on(void);
void mfft_dev_profile_bytes(double bytes);     /* algorithmic bytes of the next launch */
int  mfft_dev_profile_read(double *ms, uint64_t *launches, double *bytes, int nclass);

/* counters for bench.py: number of kernel launches issued through this ABI since reset */
uint64_t mfft_dev_launch_count(void);
void     mfft_dev_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif
