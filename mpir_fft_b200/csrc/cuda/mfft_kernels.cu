/* mfft_kernels.cu -- sm_100a kernels and the thin launch ABI declared in mfft_internal.h.
 *
 * Kernels (one warp owns one coefficient block unless noted):
 *   k_run_stage   one stage of a transform: every (op, batch entry) pair is a warp computing up to
 *                 two outputs  sA*A*2^eA + sB*B*2^eB  (butterflies / twiddles of mul_fft.c:517-957)
 *   k_finalize    gather + optional 2^shift scaling + mpn_normmod_2expp1 (mul_fft.c:272-294,
 *                 2397-2407, 3256-3260)
 *   k_pointwise   a*b mod 2^NW+1 as a 32-block negacyclic schoolbook product inside a warp
 *                 (new_mpn_mulmod_2expp1, mul_fft.c:3119-3123, 3244-3253)
 *   k_split       FFT_split_bits (mul_fft.c:115-170) + zero fill (3235-3236)
 *   k_combine_*   FFT_combine_bits (mul_fft.c:207-267) as a gather + global carry lookahead
 *
 * There is no CPU fallback anywhere in this file: without a usable device every entry point
 * returns an error and the C host side aborts with a diagnostic.
 */
#ifdef MFFT_EMU
#include "cuda_emu.h"      /* tests/emu: CPU emulation of the SIMT model, test builds only */
#else
#include <cuda_runtime.h>
#define MFFT_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MFFT_DYN_SMEM(type, name) extern __shared__ type name[]
/* launch that may overlap its scheduling and prologue with the tail of the previous kernel of the
   stream (programmatic dependent launch); the kernel must execute griddepcontrol.wait before it
   touches anything that kernel wrote */
#define MFFT_LAUNCH_PDL(on, kern, grid_, block_, smem_, stream_, ...)                                   \
   do {                                                                                            \
      if (!(on)) { kern<<<(grid_), (block_), (smem_), (stream_)>>>(__VA_ARGS__); break; }               \
      cudaLaunchConfig_t cfg__; memset(&cfg__, 0, sizeof cfg__);                                    \
      cfg__.gridDim = dim3(grid_); cfg__.blockDim = dim3(block_);                                     \
      cfg__.dynamicSmemBytes = (smem_); cfg__.stream = (stream_);                                     \
      cudaLaunchAttribute at__[1];                                                                 \
      at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                             \
      at__[0].val.programmaticStreamSerializationAllowed = 1;                                      \
      cfg__.attrs = at__; cfg__.numAttrs = 1;                                                       \
      cudaLaunchKernelEx(&cfg__, kern, __VA_ARGS__);                                                \
   } while (0)
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../mfft_internal.h"
#include "mfft_arith.h"

#define FULL 0xffffffffu
#define MFFT_PDL_DEFAULT 1
#define MFFT_COMBINE_FUSED_DEFAULT 0
#ifndef PW_UNROLL
#define PW_UNROLL 4
#endif

/* programmatic dependent launch (see MFFT_LAUNCH_PDL): first statement of a kernel that may be
   launched with the stream-serialisation attribute -- nothing the previous kernel wrote is read
   before it.  No-ops without the attribute. */
__device__ __forceinline__ void pdl_wait()
{
#ifndef MFFT_EMU
   asm volatile("griddepcontrol.wait;" ::: "memory");
   asm volatile("griddepcontrol.launch_dependents;");
#endif
}
static int g_pdl = -1;
static int pdl_on(void)
{
   if (g_pdl < 0) { const char *e = getenv("MPIRFFT_PDL"); g_pdl = e ? (e[0] != '0') : MFFT_PDL_DEFAULT; }
   return g_pdl;
}

static char g_err[512] = "";
static uint64_t g_launches = 0;

/* optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
 * leg).  Off by default; never part of the timed end-to-end path. */
enum { PC_STAGE = 0, PC_FINALIZE, PC_POINTWISE, PC_SPLIT, PC_COMBINE, PC_NORMALISE, PC_COUNT };
static int g_prof_on = 0;
static double g_prof_bytes_next = 0.0;
#ifndef MFFT_EMU
#include <vector>
struct prof_rec { cudaEvent_t a, b; int cls; double bytes; };
static std::vector<prof_rec> g_prof;
struct prof_scope {
   int on; prof_rec r; cudaStream_t st;
   prof_scope(int cls, cudaStream_t s) : on(g_prof_on), st(s)
   {
      if (!on) return;
      r.cls = cls; r.bytes = g_prof_bytes_next; g_prof_bytes_next = 0.0;
      cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st);
   }
   ~prof_scope() { if (on) { cudaEventRecord(r.b, st); g_prof.push_back(r); } }
};
#define PROF(cls, st) prof_scope prof__(cls, (cudaStream_t)(st))
#else
#define PROF(cls, st) do { } while (0)
#endif

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                       \
      snprintf(g_err, sizeof g_err, "%s:%d: %s: %s", __FILE__, __LINE__, #call,                 \
               cudaGetErrorString(e__)); return -1; } } while (0)
#define CKL() do { cudaError_t e__ = cudaGetLastError(); g_launches++; if (e__ != cudaSuccess) { \
      snprintf(g_err, sizeof g_err, "%s:%d: launch: %s", __FILE__, __LINE__,                    \
               cudaGetErrorString(e__)); return -1; } } while (0)

/* ------------------------------------------------------------------------------------------ */
/* warp-level   out = sA*A*2^eA + sB*B*2^eB   (see mfft_arith.h for the derivation)            */
/* ------------------------------------------------------------------------------------------ */
template <int M>
__device__ __forceinline__ void lincomb_out(limb_t *out, const limb_t *A, int sA, uint64_t eA,
                                            const limb_t *B, int sB, uint64_t eB, uint32_t l,
                                            uint32_t lane)
{
   int64_t top_acc = 0; int ones = 0;
   mfft_term ta, tb;
   mfft_term_setup(&ta, A, l, sA, eA, &top_acc, &ones);
   mfft_term_setup(&tb, B, l, sB, eB, &top_acc, &ones);
   uint32_t tc = ones ? 1u : 0u;
   if (ones == 2) top_acc -= 1;

   for (uint32_t k0 = 0; k0 < l; k0 += 32u * M)
   {
      const uint32_t kb = k0 + lane * M;
      limb_t s[M];
      uint32_t c = 0; bool ones_all = true;
#pragma unroll
      for (int i = 0; i < M; i++)
      {
         const uint32_t k = kb + i;
         limb_t x = ~(limb_t)0, z = 0;            /* limbs past l just pass the carry on */
         if (k < l)
         {
            x = ta.present ? mfft_term_limb(&ta, l, k) : 0;
            z = tb.present ? mfft_term_limb(&tb, l, k) : 0;
         }
         mfft_u128 acc = (mfft_u128) x + z + c;
         s[i] = (limb_t) acc; c = (uint32_t)(acc >> 64);
         ones_all = ones_all && (s[i] == ~(limb_t)0);
      }
      const uint32_t G = __ballot_sync(FULL, c != 0);
      const uint32_t P = __ballot_sync(FULL, ones_all);
      const uint64_t la = mfft_lookahead(G, P, tc);
      uint32_t myc = (uint32_t)(la >> lane) & 1u;
      tc = (uint32_t)(la >> 32) & 1u;
#pragma unroll
      for (int i = 0; i < M; i++)
      {
         const uint32_t k = kb + i;
         limb_t v = s[i] + myc;
         myc = (myc && v == 0) ? 1u : 0u;
         if (k < l) out[k] = v;
      }
   }
   top_acc += tc;
   __syncwarp();
   if (lane == 0)
   {
      if (ta.present && tb.present && ta.y == tb.y) { ta.K += tb.K; tb.K = 0; }
      if (ta.present) mfft_inject(out, l, &top_acc, ta.y, ta.K);
      if (tb.present) mfft_inject(out, l, &top_acc, tb.y, tb.K);
      out[l] = (limb_t) top_acc;
   }
   __syncwarp();
}

/* warp-level mpn_normmod_2expp1 (mul_fft.c:272-294): canonical form in [0, 2^NW] */
__device__ __forceinline__ void normalise_block(limb_t *blk, uint32_t l, uint32_t lane)
{
   for (int it = 0; it < 4; it++)
   {
      __syncwarp();
      const int64_t top = (int64_t) blk[l];
      if (top == 0) break;
      if (top == 1)
      {  /* 2^NW itself is canonical: top 1, body 0 */
         bool z = true;
         for (uint32_t k = lane; k < l; k += 32) z = z && (blk[k] == 0);
         if (__all_sync(FULL, z)) break;
      }
      __syncwarp();
      if (lane == 0)
      {
         int64_t nt = 0;
         mfft_inject(blk, l, &nt, 0, -(mfft_i128) top);
         blk[l] = (limb_t) nt;
      }
   }
   __syncwarp();
}

__device__ __forceinline__ limb_t *block_ptr(limb_t *slab, const mfft_geom &g, uint32_t slot,
                                             const mfft_batch &b)
{
   const uint32_t half = (slot / g.S) ^ b.parity;
   const uint64_t idx = (uint64_t) half * g.half_blocks + b.base + (uint64_t)(slot % g.S) * g.slot_stride;
   return slab + idx * g.pitch;
}

template <int M>
__global__ void __launch_bounds__(128)
k_run_stage(limb_t *slab, mfft_geom g, const mfft_op *__restrict__ ops, uint32_t count,
            const mfft_batch *__restrict__ batch, uint32_t nbatch)
{
   const uint64_t wid = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const uint32_t lane = threadIdx.x & 31;
   if (wid >= (uint64_t) count * nbatch) return;
   const mfft_op op = ops[wid / nbatch];
   const mfft_batch b = batch[wid % nbatch];
   const uint64_t M2 = 128ull * g.l;
   const limb_t *A = block_ptr(slab, g, op.inA, b);
   const limb_t *B = (op.inB != MFFT_NONE) ? block_ptr(slab, g, op.inB, b) : A;
   limb_t *S = block_ptr(slab, g, op.outS, b);
   lincomb_out<M>(S, A, op.sSA, (op.eSA + (uint64_t) b.col * op.cSA) % M2,
                     B, op.sSB, (op.eSB + (uint64_t) b.col * op.cSB) % M2, g.l, lane);
   if (op.outT != MFFT_NONE)
   {
      limb_t *Tt = block_ptr(slab, g, op.outT, b);
      lincomb_out<M>(Tt, A, op.sTA, (op.eTA + (uint64_t) b.col * op.cTA) % M2,
                         B, op.sTB, (op.eTB + (uint64_t) b.col * op.cTB) % M2, g.l, lane);
   }
}

template <int M>
__global__ void __launch_bounds__(128)
k_finalize(limb_t *dst, uint32_t dst_stride, const uint32_t *__restrict__ dst_base,
           limb_t *slab, mfft_geom g, const mfft_move *__restrict__ moves, uint32_t nmoves,
           const mfft_batch *__restrict__ batch, uint32_t nbatch, uint32_t shift, int normalise)
{
   const uint64_t wid = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const uint32_t lane = threadIdx.x & 31;
   if (wid >= (uint64_t) nmoves * nbatch) return;
   const mfft_move mv = moves[wid / nbatch];
   const uint32_t bi = (uint32_t)(wid % nbatch);
   const mfft_batch b = batch[bi];
   const limb_t *src = block_ptr(slab, g, mv.src_slot, b);
   limb_t *out = dst + ((uint64_t) dst_base[bi] + (uint64_t) mv.dst_pos * dst_stride) * g.pitch;
   lincomb_out<M>(out, src, 1, shift, src, 0, 0, g.l, lane);
   if (normalise) normalise_block(out, g.l, lane);
}


/* fused shared-memory tile executor (carry-save coefficients): see mfft_tiles.h */
#include "mfft_tiles.h"
/* carry-save layers on HBM-resident coefficients of any size: see mfft_cs_stage.h */
#include "mfft_cs_stage.h"

__global__ void __launch_bounds__(128)
k_normalise(limb_t *slab, uint32_t l, uint32_t pitch, uint64_t nblk)
{
   const uint64_t wid = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   if (wid >= nblk) return;
   normalise_block(slab + wid * pitch, l, threadIdx.x & 31);
}

/* ------------------------------------------------------------------------------------------ */
/* pointwise product mod 2^NW + 1                                                              */
/* ------------------------------------------------------------------------------------------ */
/* 32-bit multiply-accumulate carry chains.  ptxas fuses every mad.lo.cc / madc.hi.cc pair into one
 * IMAD.WIDE.U32(.X) with the carry in a predicate; the accumulator pair must be an aligned
 * register pair, which is why even- and odd-offset products go to separate arrays below.
 * (Emulation build: the same semantics in plain C with an explicit carry flag.) */
#ifdef MFFT_EMU
static uint32_t emu_cc;
static inline void mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b)
{ uint64_t v = (uint64_t)(uint32_t)((uint64_t) a * b) + d; d = (uint32_t) v; emu_cc = (uint32_t)(v >> 32); }
static inline void madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b)
{ uint64_t v = (uint64_t)(uint32_t)((uint64_t) a * b) + d + emu_cc; d = (uint32_t) v; emu_cc = (uint32_t)(v >> 32); }
static inline void madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b)
{ uint64_t v = (((uint64_t) a * b) >> 32) + d + emu_cc; d = (uint32_t) v; emu_cc = (uint32_t)(v >> 32); }
static inline void addc(uint32_t &d) { d += emu_cc; }
#else
__device__ __forceinline__ void mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b)
{ asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b)); }
__device__ __forceinline__ void madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b)
{ asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b)); }
__device__ __forceinline__ void madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b)
{ asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b)); }
__device__ __forceinline__ void addc(uint32_t &d)
{ asm volatile("addc.u32 %0, %0, 0;" : "+r"(d)); }
#endif

/* P += x*y for two H-word blocks (H even), P = sum e[k] 2^(32k) + sum o[k] 2^(32(k+1)) + sum kc[i] 2^(32(H+i)):
 * H rows of two IMAD.WIDE.U32.X chains (even / odd word offsets), final carries counted in kc[] */
template <int H>
__device__ __forceinline__ void mac_block(uint32_t (&e)[2 * H + 2], uint32_t (&o)[2 * H + 2], uint32_t (&kc)[H + 2],
                                          const uint32_t (&x)[H], const uint32_t (&y)[H])
{
#pragma unroll
   for (int j = 0; j < H; j++)
   {
      const uint32_t yj = y[j];
      {
         const int t0 = j & 1;
         mad_lo_cc(e[t0 + j], x[t0], yj); madc_hi_cc(e[t0 + j + 1], x[t0], yj);
#pragma unroll
         for (int t = t0 + 2; t < H; t += 2) { madc_lo_cc(e[t + j], x[t], yj); madc_hi_cc(e[t + j + 1], x[t], yj); }
         addc(kc[j + (j & 1)]);
      }
      {
         const int t0 = 1 - (j & 1);
         mad_lo_cc(o[t0 + j - 1], x[t0], yj); madc_hi_cc(o[t0 + j], x[t0], yj);
#pragma unroll
         for (int t = t0 + 2; t < H; t += 2) { madc_lo_cc(o[t + j - 1], x[t], yj); madc_hi_cc(o[t + j], x[t], yj); }
         addc(kc[j + 1 - (j & 1)]);
      }
   }
}

/* word k of the accumulator triple of mac_block, as a non-negative 64-bit value */
template <int H>
__device__ __forceinline__ int64_t mac_word(const uint32_t (&e)[2 * H + 2], const uint32_t (&o)[2 * H + 2],
                                            const uint32_t (&kc)[H + 2], int k)
{
   int64_t v = 0;
   if (k >= 0 && k < 2 * H + 2) v += (int64_t)(uint64_t) e[k];
   if (k >= 1 && k - 1 < 2 * H + 2) v += (int64_t)(uint64_t) o[k - 1];
   if (k >= H && k - H < H + 2) v += (int64_t)(uint64_t) kc[k - H];
   return v;
}

/* S_K = sum_{s > K} A_s over the lanes' C-word blocks (C+1 words): inclusive suffix scan, then shift
 * down by one lane.  Called after the product loop so that S is not live across it. */
template <int C>
__device__ __forceinline__ void suffix_block_sums(uint32_t (&S)[C + 1], const uint32_t *own, uint32_t lane)
{
#pragma unroll
   for (int i = 0; i < C; i++) S[i] = own[i];
   S[C] = 0;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1)
   {
      uint64_t cy = 0;
      const bool take = (lane + d < 32);
#pragma unroll
      for (int i = 0; i <= C; i++)
      {
         const uint32_t other = __shfl_down_sync(FULL, S[i], d);
         const uint64_t v = (uint64_t) S[i] + (take ? other : 0u) + cy;
         S[i] = (uint32_t) v; cy = v >> 32;
      }
   }
#pragma unroll
   for (int i = 0; i <= C; i++)
   {
      const uint32_t other = __shfl_down_sync(FULL, S[i], 1);
      S[i] = (lane == 31) ? 0u : other;
   }
}

/* One warp per product; l = 16*C limbs, i.e. every lane owns C 32-bit words of each operand.
 * The operands are cut into 32 blocks of C words; lane K needs
 *      sum_{I+J=K} A_I*B_J  -  sum_{I+J=K+32} A_I*B_J            (B^l == -1).
 * Step s: A_s is broadcast from shared memory, the B blocks rotate up one lane per step, so
 * lane K holds B_{(K-s) mod 32}.  A block that wraps from lane 31 to lane 0 is complemented once
 * and stays so: a*(~b) = a*2^(32C) - a - a*b, hence every term is a non-negative product and
 * the signs reduce to one correction  -S_K*2^(32C) + S_K,  S_K = sum_{s>K} A_s (a suffix sum).
 * Each CxC block product is C rows of two IMAD.WIDE.U32.X chains (even / odd word offsets,
 * accumulators e[] / o[]); a chain's final carry goes to a small counter kc[], so the chains
 * never have to ripple into words that already hold data from earlier steps.               */
/* KARA: every CxC block product is split once more inside the lane (H = C/2, W = 2^(32H)):
 *      a*b = L + (M - L - Hh) W + Hh W^2,   L = a0 b0, Hh = a1 b1, M = (a0+a1)(b0+b1),
 * and because the identity is linear, L, M, Hh are accumulated over the 32 steps in three separate
 * accumulator sets and combined ONCE after the loop: 3 H^2 instead of C^2 = 4 H^2 multiply-adds per
 * step, no subtraction inside the loop.  The carry bits of the two sums never enter a multiply:
 * (xs + ca W)(ys + cb W) = xs ys + (ca ys + cb xs) W + ca cb W^2, a masked add per step (UV, n2).
 * a0+a1 is formed once per block when the operand is staged in shared memory (H+1 more words per
 * block); b0+b1 is recomputed after every rotation (the complement of a wrapped block is not the
 * complement of its half-sum).  The additions go to the ALU pipe, next to the IMAD pipe the products
 * keep busy. */
/* One level of Karatsuba on 2Q-word operands, accumulated over the steps of a sweep (the building block of the
 * two-level variant, KARA == 2): L = x0 y0, Hh = x1 y1, M = (x0+x1)(y0+y1) in three accumulator sets, the carry
 * bits of the two sums in UV / n2 exactly as in the one-level code.  reduce() turns the sets into the 4Q+1
 * canonical words of sum_s X_s Y_s (< 32 * 2^(128 Q)). */
template <int Q>
struct kara_acc
{
   uint32_t eL[2 * Q + 2], oL[2 * Q + 2], kL[Q + 2], eM[2 * Q + 2], oM[2 * Q + 2], kM[Q + 2],
            eH[2 * Q + 2], oH[2 * Q + 2], kH[Q + 2], UV[Q + 1], n2;
   __device__ __forceinline__ void zero()
   {
#pragma unroll
      for (int i = 0; i < 2 * Q + 2; i++) { eL[i] = 0; oL[i] = 0; eM[i] = 0; oM[i] = 0; eH[i] = 0; oH[i] = 0; }
#pragma unroll
      for (int i = 0; i < Q + 2; i++) { kL[i] = 0; kM[i] = 0; kH[i] = 0; }
#pragma unroll
      for (int i = 0; i <= Q; i++) UV[i] = 0;
      n2 = 0;
   }
   /* X: x0 | x1 (2Q words, shared memory), XS: x0 + x1 (Q words) and its carry; y: y0 | y1 in registers */
   __device__ __forceinline__ void step(const uint32_t *X, const uint32_t *XS, const uint32_t (&y)[2 * Q])
   {
      uint32_t x[Q], yy[Q];
#pragma unroll
      for (int i = 0; i < Q; i++) { x[i] = X[i]; yy[i] = y[i]; }
      mac_block<Q>(eL, oL, kL, x, yy);
#pragma unroll
      for (int i = 0; i < Q; i++) { x[i] = X[Q + i]; yy[i] = y[Q + i]; }
      mac_block<Q>(eH, oH, kH, x, yy);
      uint64_t cy = 0;
#pragma unroll
      for (int i = 0; i < Q; i++)
      {
         const uint64_t v = (uint64_t) y[i] + y[Q + i] + cy;
         yy[i] = (uint32_t) v; cy = v >> 32;
         x[i] = XS[i];
      }
      const uint32_t cb = (uint32_t) cy, ca = XS[Q];
      mac_block<Q>(eM, oM, kM, x, yy);
      const uint32_t ma = 0u - ca, mb = 0u - cb;
      cy = 0;
#pragma unroll
      for (int i = 0; i < Q; i++)
      {
         const uint64_t v = (uint64_t) UV[i] + (yy[i] & ma) + (x[i] & mb) + cy;
         UV[i] = (uint32_t) v; cy = v >> 32;
      }
      UV[Q] += (uint32_t) cy;
      n2 += ca & cb;
   }
   /* V[0 .. 4Q] = L + (M + UV W + n2 W^2 - L - Hh) W + Hh W^2,  W = 2^(32 Q) */
   __device__ __forceinline__ void reduce(uint32_t *V) const
   {
      int64_t cy = 0;
#pragma unroll
      for (int k = 0; k <= 4 * Q; k++)
      {
         int64_t v = cy + mac_word<Q>(eL, oL, kL, k);
         if (k >= Q) v += mac_word<Q>(eM, oM, kM, k - Q) - mac_word<Q>(eL, oL, kL, k - Q) - mac_word<Q>(eH, oH, kH, k - Q);
         if (k >= 2 * Q) v += mac_word<Q>(eH, oH, kH, k - 2 * Q);
         if (k >= 2 * Q && k - 2 * Q <= Q) v += (int64_t)(uint64_t) UV[k - 2 * Q];
         if (k == 3 * Q) v += (int64_t)(uint64_t) n2;
         V[k] = (uint32_t) v; cy = v >> 32;
      }
   }
};

/* (Nine warps per SM -- 224 registers, three CTAs of 96 threads -- were measured: 36 % SLOWER.  The kernel is
   bound by the IMAD pipe of each scheduler, and 9 warps spread 3/2/2/2 over the four schedulers.) */
/* KARA: 0 schoolbook blocks, 1 one level of Karatsuba (two sweeps at C = 32), 2 two levels: the three
 * half-size products L, Hh, M of the first level are accumulated in sweeps of their own (L and Hh together when
 * MERGE), each with one more level inside the lane (kara_acc): 9 Q^2 = 2.25 H^2 multiply-adds per step. */
template <int C, int KARA, int UNR, bool MERGE = false>
__global__ void __launch_bounds__(128, (KARA == 2 && C == 16 && !MERGE) ? 3 : 1)      /* one form per sweep: 12 warps per SM fit */
k_pointwise(limb_t *a_slab, const limb_t *b_slab, const uint32_t *__restrict__ blocks,
            uint32_t nblk, uint32_t l, uint32_t pitch)
{
   pdl_wait();
   MFFT_DYN_SMEM(uint32_t, smem);
   constexpr int H = C / 2;
   constexpr int Q = H / 2;
   constexpr bool TWOSWEEP = (KARA == 1) && (C >= 32);          /* three accumulator sets do not fit the registers */
   constexpr int PARK = 2 * (2 * H + 1) + 1;                    /* words of L and Hh parked per lane (odd stride) */
   constexpr int WSM = KARA == 2 ? 32 * (C + H + 1 + 3 * (Q + 1) + PARK)
                     : KARA ? 32 * (C + H + 1) + (TWOSWEEP ? 32 * PARK : 0) : 32 * C;       /* shared words per warp */
   const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const uint64_t wid = (uint64_t) blockIdx.x * (blockDim.x >> 5) + wib;
   if (wid >= nblk) return;
   uint32_t *sA = smem + wib * WSM;
   uint32_t *sAs = sA + 32 * C;                                 /* KARA: a0+a1 (H words) and its carry, per block */
   limb_t *A = a_slab + (uint64_t) blocks[wid] * pitch;
   const limb_t *B = b_slab + (uint64_t) blocks[wid] * pitch;

   const int64_t topA = (int64_t) A[l], topB = (int64_t) B[l];
   __syncwarp();          /* every lane has read the tops before any lane rewrites A */
   if (topA | topB)
   {  /* canonical inputs: top == 1 means the operand is 2^NW == -1 (body zero) */
      if (topA && topB)
      {
         for (uint32_t k = lane; k < l; k += 32) A[k] = (k == 0);
         if (lane == 0) A[l] = 0;
      } else if (topA)
      {
         lincomb_out<1>(A, B, -1, 0, B, 0, 0, l, lane);
         normalise_block(A, l, lane);
      } else
      {
         __syncwarp();
         /* -A in place: each lane rewrites exactly the limbs it read in the same tile */
         lincomb_out<1>(A, A, -1, 0, A, 0, 0, l, lane);
         normalise_block(A, l, lane);
      }
      return;
   }

   uint32_t a[C], b[C];
#pragma unroll
   for (int i = 0; i < C; i += 2)
   {
      const limb_t va = A[(lane * C + i) >> 1], vb = B[(lane * C + i) >> 1];
      a[i] = (uint32_t) va; a[i + 1] = (uint32_t)(va >> 32);
      b[i] = (uint32_t) vb; b[i + 1] = (uint32_t)(vb >> 32);
   }
#pragma unroll
   for (int i = 0; i < C; i++) sA[lane * C + i] = a[i];
   if constexpr (KARA)
   {
      uint64_t cy = 0;
#pragma unroll
      for (int i = 0; i < H; i++)
      {
         const uint64_t v = (uint64_t) a[i] + a[H + i] + cy;
         sAs[lane * (H + 1) + i] = (uint32_t) v; cy = v >> 32;
      }
      sAs[lane * (H + 1) + H] = (uint32_t) cy;
      if constexpr (KARA == 2)
      {  /* second level: x0 + x1 (Q words and a carry) of the three forms a0, a1, a0 + a1 of this lane's block */
         uint32_t *sQ = sAs + 32 * (H + 1) + lane * 3 * (Q + 1);
#pragma unroll
         for (int f = 0; f < 3; f++)
         {
            uint64_t c2 = 0;
#pragma unroll
            for (int i = 0; i < Q; i++)
            {
               const uint32_t lo = f == 0 ? a[i] : f == 1 ? a[H + i] : sAs[lane * (H + 1) + i];
               const uint32_t hi = f == 0 ? a[Q + i] : f == 1 ? a[H + Q + i] : sAs[lane * (H + 1) + Q + i];
               const uint64_t v = (uint64_t) lo + hi + c2;
               sQ[f * (Q + 1) + i] = (uint32_t) v; c2 = v >> 32;
            }
            sQ[f * (Q + 1) + Q] = (uint32_t) c2;
         }
      }
   }
   __syncwarp();

   uint32_t acc[2 * C + 1];
   const uint32_t m0 = (lane == 0) ? 0xffffffffu : 0u;
   if constexpr (KARA == 2)
   {
      constexpr int NV = 2 * H + 1;                              /* words of a half-size product summed over the steps */
      const uint32_t *sQ = sAs + 32 * (H + 1);
      uint32_t *park = sAs + 32 * (H + 1) + 32 * 3 * (Q + 1) + lane * PARK;
      const uint32_t src = (lane + 31) & 31;
      if constexpr (MERGE)
      {
         kara_acc<Q> A0, A1;
         A0.zero(); A1.zero();
#pragma unroll UNR
         for (uint32_t s = 0; s < 32; s++)
         {
            uint32_t y[H];
#pragma unroll
            for (int i = 0; i < H; i++) y[i] = b[i];
            A0.step(sA + s * C, sQ + (s * 3 + 0) * (Q + 1), y);
#pragma unroll
            for (int i = 0; i < H; i++) y[i] = b[H + i];
            A1.step(sA + s * C + H, sQ + (s * 3 + 1) * (Q + 1), y);
#pragma unroll
            for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
         }
         A0.reduce(park); A1.reduce(park + NV);
      } else
      {
#pragma unroll 1
         for (int f = 0; f < 2; f++)
         {
            kara_acc<Q> A0;
            A0.zero();
#pragma unroll UNR
            for (uint32_t s = 0; s < 32; s++)
            {
               uint32_t y[H];
#pragma unroll
               for (int i = 0; i < H; i++) y[i] = f ? b[H + i] : b[i];
               A0.step(sA + s * C + f * H, sQ + (s * 3 + f) * (Q + 1), y);
#pragma unroll
               for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
            }
            A0.reduce(park + f * NV);
            /* after 32 rotations every B block has been complemented exactly once */
#pragma unroll
            for (int i = 0; i < C; i++) b[i] = ~b[i];
         }
      }
      if constexpr (MERGE)
      {
#pragma unroll
         for (int i = 0; i < C; i++) b[i] = ~b[i];
      }
      uint32_t V2[NV], UV[H + 1], n2 = 0;
      {
         kara_acc<Q> A2;
         A2.zero();
#pragma unroll
         for (int i = 0; i <= H; i++) UV[i] = 0;
#pragma unroll UNR
         for (uint32_t s = 0; s < 32; s++)
         {
            uint32_t y[H];
            uint64_t cy = 0;
#pragma unroll
            for (int i = 0; i < H; i++)
            {
               const uint64_t v = (uint64_t) b[i] + b[H + i] + cy;
               y[i] = (uint32_t) v; cy = v >> 32;
            }
            const uint32_t cb = (uint32_t) cy, ca = sAs[s * (H + 1) + H];
            A2.step(sAs + s * (H + 1), sQ + (s * 3 + 2) * (Q + 1), y);
            const uint32_t ma = 0u - ca, mb = 0u - cb;
            cy = 0;
#pragma unroll
            for (int i = 0; i < H; i++)
            {
               const uint64_t v = (uint64_t) UV[i] + (y[i] & ma) + (sAs[s * (H + 1) + i] & mb) + cy;
               UV[i] = (uint32_t) v; cy = v >> 32;
            }
            UV[H] += (uint32_t) cy;
            n2 += ca & cb;
#pragma unroll
            for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
         }
         A2.reduce(V2);
      }
      uint32_t S[C + 1];
      suffix_block_sums<C>(S, sA + lane * C, lane);
      int64_t cy = 0;
#pragma unroll
      for (int k = 0; k <= 2 * C; k++)
      {
         int64_t v = cy;
         if (k <= 2 * H) v += (int64_t)(uint64_t) park[k];
         if (k >= H && k - H <= 2 * H) v += (int64_t)(uint64_t) V2[k - H] - (int64_t)(uint64_t) park[k - H] - (int64_t)(uint64_t) park[NV + k - H];
         if (k >= C && k - C <= 2 * H) v += (int64_t)(uint64_t) park[NV + k - C];
         if (k >= C && k - C <= H) v += (int64_t)(uint64_t) UV[k - C];
         if (k == 3 * H) v += (int64_t)(uint64_t) n2;
         if (k <= C) v += (int64_t)(uint64_t) S[k];
         if (k >= C) v -= (int64_t)(uint64_t) S[k - C];
         acc[k] = (uint32_t) v; cy = v >> 32;
      }
   } else
   if constexpr (TWOSWEEP)
   {
      /* Sweep 1 accumulates L and Hh over the 32 steps, reduces them to 2H+1 canonical words each and parks
         them in shared memory; after 32 rotations every B block has been complemented exactly once, so one
         more complement restores B.  Sweep 2 accumulates M (and the carry terms UV, n2) and combines. */
      uint32_t *park = sAs + 32 * (H + 1) + lane * PARK;
      {
         uint32_t eL[2 * H + 2], oL[2 * H + 2], kL[H + 2], eH[2 * H + 2], oH[2 * H + 2], kH[H + 2];
#pragma unroll
         for (int i = 0; i < 2 * H + 2; i++) { eL[i] = 0; oL[i] = 0; eH[i] = 0; oH[i] = 0; }
#pragma unroll
         for (int i = 0; i < H + 2; i++) { kL[i] = 0; kH[i] = 0; }
#pragma unroll UNR
         for (uint32_t s = 0; s < 32; s++)
         {
            uint32_t x[H], y[H];
#pragma unroll
            for (int i = 0; i < H; i++) { x[i] = sA[s * C + i]; y[i] = b[i]; }
            mac_block<H>(eL, oL, kL, x, y);
#pragma unroll
            for (int i = 0; i < H; i++) { x[i] = sA[s * C + H + i]; y[i] = b[H + i]; }
            mac_block<H>(eH, oH, kH, x, y);
            const uint32_t src = (lane + 31) & 31;
#pragma unroll
            for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
         }
         int64_t cl = 0, ch = 0;
#pragma unroll
         for (int k = 0; k <= 2 * H; k++)
         {
            const int64_t vl = cl + mac_word<H>(eL, oL, kL, k), vh = ch + mac_word<H>(eH, oH, kH, k);
            park[k] = (uint32_t) vl; cl = vl >> 32;
            park[2 * H + 1 + k] = (uint32_t) vh; ch = vh >> 32;
         }
      }
#pragma unroll
      for (int i = 0; i < C; i++) b[i] = ~b[i];
      uint32_t eM[2 * H + 2], oM[2 * H + 2], kM[H + 2], UV[H + 1], n2 = 0;
#pragma unroll
      for (int i = 0; i < 2 * H + 2; i++) { eM[i] = 0; oM[i] = 0; }
#pragma unroll
      for (int i = 0; i < H + 2; i++) kM[i] = 0;
#pragma unroll
      for (int i = 0; i <= H; i++) UV[i] = 0;
#pragma unroll UNR
      for (uint32_t s = 0; s < 32; s++)
      {
         uint32_t x[H], y[H];
         uint64_t cy = 0;
#pragma unroll
         for (int i = 0; i < H; i++)
         {
            const uint64_t v = (uint64_t) b[i] + b[H + i] + cy;
            y[i] = (uint32_t) v; cy = v >> 32;
            x[i] = sAs[s * (H + 1) + i];
         }
         const uint32_t cb = (uint32_t) cy, ca = sAs[s * (H + 1) + H];
         mac_block<H>(eM, oM, kM, x, y);
         const uint32_t ma = 0u - ca, mb = 0u - cb;
         cy = 0;
#pragma unroll
         for (int i = 0; i < H; i++)
         {
            const uint64_t v = (uint64_t) UV[i] + (y[i] & ma) + (x[i] & mb) + cy;
            UV[i] = (uint32_t) v; cy = v >> 32;
         }
         UV[H] += (uint32_t) cy;
         n2 += ca & cb;
         const uint32_t src = (lane + 31) & 31;
#pragma unroll
         for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
      }
      uint32_t S[C + 1];
      suffix_block_sums<C>(S, sA + lane * C, lane);
      int64_t cy = 0;
#pragma unroll
      for (int k = 0; k <= 2 * C; k++)
      {
         int64_t v = cy;
         if (k <= 2 * H) v += (int64_t)(uint64_t) park[k];
         if (k >= H) v += mac_word<H>(eM, oM, kM, k - H);
         if (k >= H && k - H <= 2 * H) v -= (int64_t)(uint64_t) park[k - H] + (int64_t)(uint64_t) park[2 * H + 1 + k - H];
         if (k >= C && k - C <= 2 * H) v += (int64_t)(uint64_t) park[2 * H + 1 + k - C];
         if (k >= C && k - C <= H) v += (int64_t)(uint64_t) UV[k - C];
         if (k == 3 * H) v += (int64_t)(uint64_t) n2;
         if (k <= C) v += (int64_t)(uint64_t) S[k];
         if (k >= C) v -= (int64_t)(uint64_t) S[k - C];
         acc[k] = (uint32_t) v; cy = v >> 32;
      }
   } else if constexpr (KARA)
   {
      uint32_t eL[2 * H + 2], oL[2 * H + 2], kL[H + 2], eM[2 * H + 2], oM[2 * H + 2], kM[H + 2],
               eH[2 * H + 2], oH[2 * H + 2], kH[H + 2], UV[H + 1], n2 = 0;
#pragma unroll
      for (int i = 0; i < 2 * H + 2; i++) { eL[i] = 0; oL[i] = 0; eM[i] = 0; oM[i] = 0; eH[i] = 0; oH[i] = 0; }
#pragma unroll
      for (int i = 0; i < H + 2; i++) { kL[i] = 0; kM[i] = 0; kH[i] = 0; }
#pragma unroll
      for (int i = 0; i <= H; i++) UV[i] = 0;

#pragma unroll UNR
      for (uint32_t s = 0; s < 32; s++)
      {
         uint32_t x[H], y[H];
#pragma unroll
         for (int i = 0; i < H; i++) { x[i] = sA[s * C + i]; y[i] = b[i]; }
         mac_block<H>(eL, oL, kL, x, y);
#pragma unroll
         for (int i = 0; i < H; i++) { x[i] = sA[s * C + H + i]; y[i] = b[H + i]; }
         mac_block<H>(eH, oH, kH, x, y);
         uint64_t cy = 0;
#pragma unroll
         for (int i = 0; i < H; i++)
         {
            const uint64_t v = (uint64_t) b[i] + b[H + i] + cy;
            y[i] = (uint32_t) v; cy = v >> 32;
            x[i] = sAs[s * (H + 1) + i];
         }
         const uint32_t cb = (uint32_t) cy, ca = sAs[s * (H + 1) + H];
         mac_block<H>(eM, oM, kM, x, y);
         /* carry bits of the two sums: UV += ca*ys + cb*xs, n2 += ca*cb */
         const uint32_t ma = 0u - ca, mb = 0u - cb;
         cy = 0;
#pragma unroll
         for (int i = 0; i < H; i++)
         {
            const uint64_t v = (uint64_t) UV[i] + (y[i] & ma) + (x[i] & mb) + cy;
            UV[i] = (uint32_t) v; cy = v >> 32;
         }
         UV[H] += (uint32_t) cy;
         n2 += ca & cb;
         const uint32_t src = (lane + 31) & 31;
#pragma unroll
         for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
      }

      /* P_K = L + (M + UV W + n2 W^2 - L - Hh) W + Hh W^2;   R_K = P_K - S_K*2^(32C) + S_K */
      uint32_t S[C + 1];
      suffix_block_sums<C>(S, sA + lane * C, lane);
      int64_t cy = 0;
#pragma unroll
      for (int k = 0; k <= 2 * C; k++)
      {
         int64_t v = cy + mac_word<H>(eL, oL, kL, k);
         if (k >= H) v += mac_word<H>(eM, oM, kM, k - H) - mac_word<H>(eL, oL, kL, k - H) - mac_word<H>(eH, oH, kH, k - H);
         if (k >= C) v += mac_word<H>(eH, oH, kH, k - C);
         if (k >= C && k - C <= H) v += (int64_t)(uint64_t) UV[k - C];
         if (k == 3 * H) v += (int64_t)(uint64_t) n2;
         if (k <= C) v += (int64_t)(uint64_t) S[k];
         if (k >= C) v -= (int64_t)(uint64_t) S[k - C];
         acc[k] = (uint32_t) v; cy = v >> 32;
      }
   } else
   {
   uint32_t e[2 * C + 2], o[2 * C + 2], kc[C + 2];
#pragma unroll
   for (int i = 0; i < 2 * C + 2; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
   for (int i = 0; i < C + 2; i++) kc[i] = 0;

#pragma unroll UNR
   for (uint32_t s = 0; s < 32; s++)
   {
      uint32_t ab[C];
#pragma unroll
      for (int i = 0; i < C; i++) ab[i] = sA[s * C + i];
#pragma unroll
      for (int j = 0; j < C; j++)
      {
         const uint32_t bj = b[j];
         /* even offsets t + j: t = (j & 1), +2, ... -> e[]; chain end at word C + j + (j & 1) */
         {
            const int t0 = j & 1;
            mad_lo_cc(e[t0 + j], ab[t0], bj); madc_hi_cc(e[t0 + j + 1], ab[t0], bj);
#pragma unroll
            for (int t = t0 + 2; t < C; t += 2) { madc_lo_cc(e[t + j], ab[t], bj); madc_hi_cc(e[t + j + 1], ab[t], bj); }
            addc(kc[j + (j & 1)]);
         }
         /* odd offsets: t = 1 - (j & 1), +2, ... -> o[]; chain end at word C + j + 1 - (j & 1) */
         {
            const int t0 = 1 - (j & 1);          /* o[i] holds word offset i + 1, so pairs start at even i */
            mad_lo_cc(o[t0 + j - 1], ab[t0], bj); madc_hi_cc(o[t0 + j], ab[t0], bj);
#pragma unroll
            for (int t = t0 + 2; t < C; t += 2) { madc_lo_cc(o[t + j - 1], ab[t], bj); madc_hi_cc(o[t + j], ab[t], bj); }
            addc(kc[j + 1 - (j & 1)]);
         }
      }
      /* rotate the B blocks up one lane; the block entering lane 0 wrapped: complement it */
      const uint32_t src = (lane + 31) & 31;
#pragma unroll
      for (int i = 0; i < C; i++) b[i] = __shfl_sync(FULL, b[i], src) ^ m0;
   }

   /* R_K = P_K - S_K*2^(32C) + S_K with P_K = e + o + kc<<(32C), as 2C words + a small signed top */
   uint32_t S[C + 1];
   suffix_block_sums<C>(S, sA + lane * C, lane);
   {
      int64_t cy = 0;
#pragma unroll
      for (int k = 0; k < 2 * C; k++)
      {
         int64_t v = cy + (int64_t)(uint64_t) e[k] + (k ? (int64_t)(uint64_t) o[k - 1] : 0);
         if (k >= C) v += (int64_t)(uint64_t) kc[k - C] - (int64_t)(uint64_t) S[k - C];
         if (k <= C) v += (int64_t)(uint64_t) S[k];
         acc[k] = (uint32_t) v; cy = v >> 32;
      }
      cy += (int64_t)(uint64_t) e[2 * C] + (int64_t)(uint64_t) o[2 * C - 1] + (int64_t)(uint64_t) kc[C] - (int64_t)(uint64_t) S[C];
      acc[2 * C] = (uint32_t)(int32_t) cy;
   }
   }

   /* lane K: r = acc[0..C) + (high part of lane K-1), lane 0 takes -(high part of lane 31) */
   uint32_t hi[C + 1];
   const uint32_t src = (lane + 31) & 31;
#pragma unroll
   for (int i = 0; i <= C; i++) hi[i] = __shfl_sync(FULL, acc[C + i], src);
   int32_t carry;
   {
      const uint32_t m = (lane == 0) ? 0xffffffffu : 0u;
      uint64_t cy = m & 1u;
#pragma unroll
      for (int i = 0; i < C; i++)
      {
         const uint64_t v = (uint64_t) acc[i] + (hi[i] ^ m) + cy;
         acc[i] = (uint32_t) v; cy = v >> 32;
      }
      carry = (int32_t)((int64_t)(int32_t)(hi[C] ^ m) + (int64_t) cy);
   }
   /* signed carries ripple to the next lane; lane 31's leave through the top limb (2^NW == -1) */
   int64_t top = 0;
   for (int it = 0; it < 40; it++)
   {
      if (!__any_sync(FULL, carry != 0)) break;
      int32_t cin = __shfl_up_sync(FULL, carry, 1);
      const int32_t c31 = __shfl_sync(FULL, carry, 31);
      top += c31;
      if (lane == 0) cin = 0;
      int64_t cc = cin;
#pragma unroll
      for (int i = 0; i < C; i++)
      {
         const int64_t v = (int64_t)(uint64_t) acc[i] + cc;
         acc[i] = (uint32_t) v; cc = v >> 32;
      }
      carry = (int32_t) cc;
   }
#pragma unroll
   for (int i = 0; i < C; i += 2)
      A[(lane * C + i) >> 1] = (limb_t) acc[i] | ((limb_t) acc[i + 1] << 32);
   if (lane == 0) A[l] = (limb_t) top;
   normalise_block(A, l, lane);
}

/* ------------------------------------------------------------------------------------------ */
/* pointwise product by a nested Schoenhage-Strassen step inside one warp                       */
/* ------------------------------------------------------------------------------------------ */
/* a*b mod 2^(64L)+1 as a negacyclic convolution of 2n' pieces of 64*PL bits over the small ring
 * p' = 2^(64*LP)+1 (the scheme of fft_mulmod_2expp1 / FFT_mulmod_2expp1, mul_fft.c:3125-3167,
 * 2998-3117, with the inner ring chosen wide enough -- 64*LP >= 2*64*PL + log2(2n') + 1 -- that the
 * signed convolution coefficients are recovered exactly and the reference's one-limb CRT fix-up
 * (2981-2996, 3067-3081) is not needed).  At L = 256 this replaces 65 536 limb products by
 * 64 * 81 = 5 184.  The inner transforms are tiny (2n' coefficients of LP limbs), so one warp
 * owns a whole product: coefficients live in shared memory (odd pitch: conflict-free), ONE LANE
 * executes one inner butterfly with serial in-lane carries -- no ballots, no cross-lane carries.
 * 2^w' (w' = 64*LP/n', even) is the 2n'-th root of unity, theta = 2^(w'/2) the negacyclic weight. */

/* limb k of (+-)X*2^e for a coefficient in shared memory, e = 64y+bs < 64*LP, complemented in the
 * negated bit region (same algebra as mfft_arith.h) */
template <int LP>
__device__ __forceinline__ limb_t ss_limb(const limb_t *X, uint32_t k, uint32_t y, uint32_t bs, uint32_t negm)
{
   uint32_t q = k + LP - y;
   if (q >= (uint32_t) LP) q -= LP;
   limb_t v = X[q];
   if (bs)
   {
      const uint32_t q1 = q ? q - 1 : LP - 1;
      v = (v << bs) | (X[q1] >> (64 - bs));
   }
   const uint32_t m32 = (uint32_t)((int32_t)(k - y) >> 31) ^ negm;
   v ^= ((limb_t) m32 << 32) | m32;
   if (k == y) v ^= (((limb_t) 1 << bs) - 1);
   return v;
}

/* out = sa*A*2^ea + sb*B*2^eb (mod p'), serial in one lane; exponents mod 2*64*LP; sb == 0: unary */
template <int LP>
__device__ __forceinline__ void ss_lincomb(limb_t (&out)[LP], int64_t &top_out, const limb_t *A, int sa, uint32_t ea,
                                           const limb_t *B, int sb, uint32_t eb)
{
   constexpr uint32_t NW = 64u * LP;
   int64_t top_acc = 0; mfft_i128 carry = 0, Ka = 0, Kb = 0;
   uint32_t ya = 0, bsa = 0, yb = 0, bsb = 0, na = 0, nb = 0;
   {
      if (ea >= NW) { ea -= NW; sa = -sa; }
      ya = ea >> 6; bsa = ea & 63; na = (sa < 0) ? 0xffffffffu : 0u;
      const int64_t top = (int64_t) A[LP];
      if (ea == 0) { if (sa > 0) top_acc += top; else { carry += 1; top_acc -= 1 + top; } }
      else if (sa > 0) { carry += 1; Ka = -(((mfft_i128)(1 + top)) << bsa); }
      else { top_acc -= 1; Ka = ((mfft_i128)(1 + top)) << bsa; }
   }
   if (sb)
   {
      if (eb >= NW) { eb -= NW; sb = -sb; }
      yb = eb >> 6; bsb = eb & 63; nb = (sb < 0) ? 0xffffffffu : 0u;
      const int64_t top = (int64_t) B[LP];
      if (eb == 0) { if (sb > 0) top_acc += top; else { carry += 1; top_acc -= 1 + top; } }
      else if (sb > 0) { carry += 1; Kb = -(((mfft_i128)(1 + top)) << bsb); }
      else { top_acc -= 1; Kb = ((mfft_i128)(1 + top)) << bsb; }
   }
#pragma unroll
   for (int k = 0; k < LP; k++)
   {
      mfft_i128 acc = carry + (mfft_i128)(mfft_u128) ss_limb<LP>(A, k, ya, bsa, na);
      if (sb) acc += (mfft_i128)(mfft_u128) ss_limb<LP>(B, k, yb, bsb, nb);
      if ((uint32_t) k == ya) acc += Ka;
      if (sb && (uint32_t) k == yb) acc += Kb;
      out[k] = (limb_t) acc; carry = acc >> 64;
   }
   top_out = top_acc + (int64_t) carry;
}

/* canonical form in one lane: body in r[], top in/out */
template <int LP>
__device__ __forceinline__ void ss_normalise(limb_t (&r)[LP], int64_t &top)
{
   for (int it = 0; it < 4; it++)
   {
      if (top == 0) return;
      if (top == 1)
      {
         limb_t any = 0;
#pragma unroll
         for (int k = 0; k < LP; k++) any |= r[k];
         if (!any) return;
      }
      mfft_i128 c = -(mfft_i128) top; top = 0;
#pragma unroll
      for (int k = 0; k < LP; k++)
      {
         const mfft_i128 acc = c + (mfft_i128)(mfft_u128) r[k];
         r[k] = (limb_t) acc; c = acc >> 64;
      }
      top = (int64_t) c;
   }
}

template <int LP>
__device__ __forceinline__ void ss_store(limb_t *X, const limb_t (&r)[LP], int64_t top)
{
#pragma unroll
   for (int k = 0; k < LP; k++) X[k] = r[k];
   X[LP] = (limb_t) top;
}

/* r = a*b mod p' for canonical a, b (registers), canonical result */
template <int LP>
__device__ __forceinline__ void ss_mulmod(limb_t (&r)[LP], int64_t &top, const limb_t (&a)[LP], int64_t ta,
                                          const limb_t (&b)[LP], int64_t tb)
{
   if (ta | tb)
   {  /* 2^NW' == -1 */
      if (ta && tb) { for (int k = 0; k < LP; k++) r[k] = (k == 0); top = 0; return; }
      mfft_i128 c = 1;
#pragma unroll
      for (int k = 0; k < LP; k++)
      {
         const mfft_i128 acc = c + (mfft_i128)(mfft_u128)(~(ta ? b[k] : a[k]));
         r[k] = (limb_t) acc; c = acc >> 64;
      }
      top = (int64_t) c - 1;            /* -x = ~x + 1 - 2^NW' */
      ss_normalise<LP>(r, top);
      return;
   }
   limb_t lo[LP], hi[LP];
   mfft_u128 acc = 0; uint32_t ovf = 0;
#pragma unroll
   for (int c = 0; c < 2 * LP - 1; c++)
   {
#pragma unroll
      for (int i = 0; i < LP; i++)
      {
         const int j = c - i;
         if (j < 0 || j >= LP) continue;
         const mfft_u128 pr = (mfft_u128) a[i] * b[j];
         const mfft_u128 old = acc; acc += pr; ovf += (acc < old);
      }
      if (c < LP) lo[c] = (limb_t) acc; else hi[c - LP] = (limb_t) acc;
      acc = (acc >> 64) | ((mfft_u128) ovf << 64); ovf = 0;
   }
   hi[LP - 1] = (limb_t) acc;
   /* lo - hi, borrow -> top */
   mfft_i128 c = 0;
#pragma unroll
   for (int k = 0; k < LP; k++)
   {
      const mfft_i128 v = c + (mfft_i128)(mfft_u128) lo[k] - (mfft_i128)(mfft_u128) hi[k];
      r[k] = (limb_t) v; c = v >> 64;
   }
   top = (int64_t) c;
   ss_normalise<LP>(r, top);
}

template <int LP>
__global__ void __launch_bounds__(128)
k_mulmod_ss(limb_t *a_slab, const limb_t *b_slab, const uint32_t *__restrict__ blocks, uint32_t nblk,
            uint32_t L, uint32_t pitch, uint32_t np, uint32_t wp, uint32_t depthp)
{
   MFFT_DYN_SMEM(limb_t, smem64);
   constexpr uint32_t CP = (LP + 1) | 1;                 /* odd pitch >= LP+1: body + signed top */
   constexpr uint32_t NW = 64u * LP, M2 = 2u * NW;
   const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const uint64_t wid = (uint64_t) blockIdx.x * (blockDim.x >> 5) + wib;
   if (wid >= nblk) return;
   const uint32_t ncoef = 2 * np, PL = L / ncoef;        /* limbs per piece */
   limb_t *ca = smem64 + (size_t) wib * 2 * ncoef * CP, *cb = ca + (size_t) ncoef * CP;
   limb_t *A = a_slab + (uint64_t) blocks[wid] * pitch;
   const limb_t *B = b_slab + (uint64_t) blocks[wid] * pitch;

   const int64_t topA = (int64_t) A[L], topB = (int64_t) B[L];
   __syncwarp();
   if (topA | topB)
   {  /* an operand equal to 2^(64L) == -1 */
      if (topA && topB) { for (uint32_t k = lane; k < L; k += 32) A[k] = (k == 0); if (lane == 0) A[L] = 0; }
      else if (topA) { lincomb_out<1>(A, B, -1, 0, B, 0, 0, L, lane); normalise_block(A, L, lane); }
      else { __syncwarp(); lincomb_out<1>(A, A, -1, 0, A, 0, 0, L, lane); normalise_block(A, L, lane); }
      return;
   }

   /* split into pieces (FFT_split, mul_fft.c:87-106; pieces are limb aligned) */
   for (uint32_t idx = lane; idx < ncoef * CP; idx += 32)
   {
      const uint32_t i = idx / CP, k = idx % CP;
      ca[idx] = (k < PL) ? A[i * PL + k] : 0;
      cb[idx] = (k < PL) ? B[i * PL + k] : 0;
   }
   __syncwarp();

   /* forward negacyclic transforms of both operands (FFT_radix2_negacyclic, 1290-1390, even w):
      first layer fuses the weights theta^i; then plain DIF layers */
   const uint32_t h = wp / 2;
   for (uint32_t t = lane; t < 2 * np; t += 32)
   {
      limb_t *X = (t < np) ? ca : cb;
      const uint32_t i = (t < np) ? t : t - np;
      limb_t *P = X + (size_t) i * CP, *Q = X + (size_t)(i + np) * CP;
      const uint32_t ea = (i * h) % M2, eb = ((i + np) * h) % M2, e = (i * wp) % M2;
      limb_t rs[LP], rt[LP]; int64_t ts, tt;
      ss_lincomb<LP>(rs, ts, P, 1, ea, Q, 1, eb);
      ss_lincomb<LP>(rt, tt, P, 1, (ea + e) % M2, Q, -1, (eb + e) % M2);
      ss_store<LP>(P, rs, ts); ss_store<LP>(Q, rt, tt);
   }
   __syncwarp();
   for (uint32_t half = np / 2, ww = 2 * wp; half >= 1; half >>= 1, ww *= 2)
   {
      for (uint32_t t = lane; t < 2 * np; t += 32)
      {
         limb_t *X = (t < np) ? ca : cb;
         const uint32_t u = (t < np) ? t : t - np;
         const uint32_t blk = u / half, i = u % half;
         limb_t *P = X + (size_t)(blk * 2 * half + i) * CP, *Q = P + (size_t) half * CP;
         const uint32_t e = (uint32_t)(((uint64_t) i * ww) % M2);
         limb_t rs[LP], rt[LP]; int64_t ts, tt;
         ss_lincomb<LP>(rs, ts, P, 1, 0, Q, 1, 0);
         ss_lincomb<LP>(rt, tt, P, 1, e, Q, -1, e);
         ss_store<LP>(P, rs, ts); ss_store<LP>(Q, rt, tt);
      }
      __syncwarp();
   }

   /* pointwise products in the small ring (outputs of the DIF transforms are in the same
      bit-reversed order for both operands, so index-wise products are what is needed) */
   for (uint32_t i = lane; i < ncoef; i += 32)
   {
      limb_t x[LP], y[LP], r[LP]; int64_t tx, ty, tr;
      limb_t *P = ca + (size_t) i * CP, *Q = cb + (size_t) i * CP;
#pragma unroll
      for (int k = 0; k < LP; k++) { x[k] = P[k]; y[k] = Q[k]; }
      tx = (int64_t) P[LP]; ty = (int64_t) Q[LP];
      ss_normalise<LP>(x, tx); ss_normalise<LP>(y, ty);
      ss_mulmod<LP>(r, tr, x, tx, y, ty);
      ss_store<LP>(P, r, tr);
   }
   __syncwarp();

   /* inverse transform (IFFT_radix2_negacyclic, 1861-1962): DIT layers, last layer fuses theta^-i */
   for (uint32_t half = 1, ww = (np > 1) ? wp * np : wp; half <= np / 2; half <<= 1, ww /= 2)
   {
      for (uint32_t t = lane; t < np; t += 32)
      {
         const uint32_t blk = t / half, i = t % half;
         limb_t *P = ca + (size_t)(blk * 2 * half + i) * CP, *Q = P + (size_t) half * CP;
         const uint32_t e = (uint32_t)((M2 - ((uint64_t) i * ww) % M2) % M2);
         limb_t rs[LP], rt[LP]; int64_t ts, tt;
         ss_lincomb<LP>(rs, ts, P, 1, 0, Q, 1, e);
         ss_lincomb<LP>(rt, tt, P, 1, 0, Q, -1, e);
         ss_store<LP>(P, rs, ts); ss_store<LP>(Q, rt, tt);
      }
      __syncwarp();
   }
   {
      /* last layer (pairs i, i+n') + weights theta^-i, theta^-(i+n') + 1/(2n') = 2^-(depth'+1) */
      const uint32_t sc = M2 - (depthp + 1);
      for (uint32_t i = lane; i < np; i += 32)
      {
         limb_t *P = ca + (size_t) i * CP, *Q = ca + (size_t)(i + np) * CP;
         const uint32_t e = (M2 - (i * wp) % M2) % M2;
         const uint32_t ua = (M2 - (i * h) % M2 + sc) % M2, ub = (M2 - ((i + np) * h) % M2 + sc) % M2;
         limb_t rs[LP], rt[LP]; int64_t ts, tt;
         ss_lincomb<LP>(rs, ts, P, 1, ua, Q, 1, (ua + e) % M2);
         ss_lincomb<LP>(rt, tt, P, 1, ub, Q, -1, (ub + e) % M2);
         ss_normalise<LP>(rs, ts); ss_normalise<LP>(rt, tt);
         ss_store<LP>(P, rs, ts); ss_store<LP>(Q, rt, tt);
      }
   }
   __syncwarp();

   /* recombine: sum_i c_i * 2^(64*PL*i) with c_i the SIGNED representative (residues above p'/2
      are negative: c = res - 2^NW' - 1); pieces at or past limb L wrap around negated.
      Lane owns output limbs [lane*G, lane*G + G), G = L/32; signed carries then ripple lane to lane. */
   const uint32_t G = L / 32;
   int64_t carry_out = 0;
   for (uint32_t g = 0; g < G; g++)
   {
      const uint32_t k = lane * G + g;
      mfft_i128 acc = carry_out;
      /* coefficients covering limb k (mod L): offsets i*PL <= pos < i*PL + LP + 1 for pos = k and pos = k + L */
      for (uint32_t wrap = 0; wrap < 2; wrap++)
      {
         const uint32_t pos = k + wrap * L;
         const int32_t ihi = (int32_t)(pos / PL);
         int32_t ilo = (int32_t)((pos >= (uint32_t) LP) ? (pos - LP) / PL : 0);
         for (int32_t i = ilo; i <= ihi; i++)
         {
            if (i < 0 || (uint32_t) i >= ncoef) continue;
            const uint32_t off = pos - (uint32_t) i * PL;
            if (off > (uint32_t) LP) continue;
            const limb_t *Cc = ca + (size_t) i * CP;
            /* sign of c_i: negative iff top == 1 or the top bit of the body is set */
            const bool negc = (Cc[LP] != 0) || (Cc[LP - 1] >> 63);
            mfft_i128 v = 0;
            if (off < (uint32_t) LP) v = (mfft_i128)(mfft_u128) Cc[off];
            else v = (mfft_i128)(int64_t) Cc[LP];                       /* top limb (0 or 1) at offset LP */
            if (negc) { if (off == 0) v -= 1; if (off == (uint32_t) LP) v -= 1; }
            acc += wrap ? -v : v;
         }
      }
      cb[k] = (limb_t) acc;   /* reuse cb as output */
      carry_out = (int64_t)(acc >> 64);
   }
   /* negative coefficients also owe their sign extension past offset LP: c_i = body - 2^NW' - 1 was
      handled above with the two -1; nothing else.  Now ripple the signed lane carries. */
   limb_t *outp = cb;
   int64_t top = 0;
   __syncwarp();
   {
      int64_t carry = carry_out;
      for (int it = 0; it < 40; it++)
      {
         if (!__any_sync(FULL, carry != 0)) break;
         int64_t cin = __shfl_up_sync(FULL, carry, 1);
         const int64_t c31 = __shfl_sync(FULL, carry, 31);
         top += c31;
         if (lane == 0) cin = 0;
         mfft_i128 cc = cin;
         for (uint32_t g = 0; g < G && cc != 0; g++)
         {
            const mfft_i128 v = cc + (mfft_i128)(mfft_u128) outp[lane * G + g];
            outp[lane * G + g] = (limb_t) v; cc = v >> 64;
         }
         carry = (int64_t) cc;
      }
   }
   __syncwarp();
   for (uint32_t k = lane; k < L; k += 32) A[k] = outp[k];
   if (lane == 0) A[L] = (limb_t) top;
   normalise_block(A, L, lane);
}

/* generic fallback: any l.  One CTA per product, thread per 64-bit output column. */
__global__ void __launch_bounds__(256)
k_pointwise_generic(limb_t *a_slab, const limb_t *b_slab, const uint32_t *__restrict__ blocks,
                    uint32_t nblk, uint32_t l, uint32_t pitch, limb_t *scratch)
{
   MFFT_DYN_SMEM(uint32_t, smem);
   limb_t *sa = (limb_t *) smem, *sb = sa + l;
   limb_t *A = a_slab + (uint64_t) blocks[blockIdx.x] * pitch;
   const limb_t *B = b_slab + (uint64_t) blocks[blockIdx.x] * pitch;
   limb_t *col = scratch + (uint64_t) blockIdx.x * 3 * l;     /* 3 limbs per column: lo, hi, signed ovf */
   const int64_t topA = (int64_t) A[l], topB = (int64_t) B[l];
   const uint32_t lane = threadIdx.x & 31;
   __syncthreads();       /* every thread has read the tops before warp 0 rewrites A */
   if (topA | topB)
   {
      if (threadIdx.x < 32)
      {
         if (topA && topB) { for (uint32_t k = lane; k < l; k += 32) A[k] = (k == 0); if (lane == 0) A[l] = 0; }
         else if (topA) { lincomb_out<1>(A, B, -1, 0, B, 0, 0, l, lane); normalise_block(A, l, lane); }
         else { lincomb_out<1>(A, A, -1, 0, A, 0, 0, l, lane); normalise_block(A, l, lane); }
      }
      return;
   }
   for (uint32_t k = threadIdx.x; k < l; k += blockDim.x) { sa[k] = A[k]; sb[k] = B[k]; }
   __syncthreads();
   for (uint32_t c = threadIdx.x; c < l; c += blockDim.x)
   {
      mfft_u128 pos = 0, neg = 0; uint32_t pov = 0, nov = 0;
      for (uint32_t i = 0; i < l; i++)
      {
         const uint32_t j = (c >= i) ? c - i : c + l - i;
         const mfft_u128 pr = (mfft_u128) sa[i] * sb[j];
         if (c >= i) { const mfft_u128 o = pos; pos += pr; pov += (pos < o); }
         else        { const mfft_u128 o = neg; neg += pr; nov += (neg < o); }
      }
      /* column value = pos - neg as a signed 192-bit number */
      const mfft_u128 d = pos - neg;
      const int64_t ov = (int64_t) pov - (int64_t) nov - (pos < neg ? 1 : 0);
      col[3 * c] = (limb_t) d; col[3 * c + 1] = (limb_t)(d >> 64); col[3 * c + 2] = (limb_t) ov;
   }
   __syncthreads();
   if (threadIdx.x == 0)
   {  /* serial carry pass (fallback path only).  Column c holds lo at limb c, hi at c+1 and a
         signed overflow count at c+2; limbs l, l+1, l+2 of the sum wrap around negated. */
      mfft_i128 cy = 0; limb_t hi3[3]; int64_t top = 0;
      for (uint32_t k = 0; k < l + 2; k++)
      {
         mfft_i128 v = cy;
         if (k < l) v += (mfft_i128)(mfft_u128) col[3 * k];
         if (k >= 1 && k - 1 < l) v += (mfft_i128)(mfft_u128) col[3 * (k - 1) + 1];
         if (k >= 2 && k - 2 < l) v += (mfft_i128)(int64_t) col[3 * (k - 2) + 2];
         if (k < l) A[k] = (limb_t) v; else hi3[k - l] = (limb_t) v;
         cy = v >> 64;
      }
      hi3[2] = (limb_t) cy;                      /* signed */
      A[l] = 0;
      for (uint32_t t = 0; t < 3; t++)
      {  /* add hi3[t] * B^(l+t):  B^l == -1, so position (l+t) mod l with the sign flipped once
            per wrap */
         mfft_i128 val = (t == 2) ? (mfft_i128)(int64_t) hi3[t] : (mfft_i128)(mfft_u128) hi3[t];
         uint32_t pos = l + t;
         while (pos >= l) { pos -= l; val = -val; }
         mfft_inject(A, l, &top, pos, val);
      }
      A[l] = (limb_t) top;
   }
   __syncthreads();
   if (threadIdx.x < 32) normalise_block(A, l, lane);
}

/* ------------------------------------------------------------------------------------------ */
/* split / combine                                                                             */
/* ------------------------------------------------------------------------------------------ */
/* bits [off, off+64) of {src, n} (zero beyond the end) */
__device__ __forceinline__ limb_t bits_at(const limb_t *src, uint64_t n, uint64_t off)
{
   const uint64_t q = off >> 6; const uint32_t r = (uint32_t)(off & 63);
   if (q >= n) return 0;
   limb_t v = src[q] >> r;
   if (r && q + 1 < n) v |= src[q + 1] << (64 - r);
   return v;
}

/* Coefficient index of local block i: the blocks of a rank's column shard are rows of `ncl`
 * consecutive coefficients taken every `n1` (ncl == n1 == 0: plain split, block i is coefficient i) */
__global__ void __launch_bounds__(256)
k_split(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *__restrict__ src, uint64_t nlimbs,
        uint64_t bits, uint64_t ncoef, uint64_t nblocks, uint32_t ncl, uint32_t n1, uint32_t c0)
{
   const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
   const uint64_t per = (uint64_t) l + 1;
   if (t >= nblocks * per) return;
   const uint64_t blk = t / per; const uint32_t k = (uint32_t)(t % per);
   const uint64_t i = ncl ? (blk / ncl) * n1 + c0 + (blk % ncl) : blk;
   limb_t v = 0;
   if (i < ncoef && k < l && (uint64_t) k * 64 < bits)
   {
      v = bits_at(src, nlimbs, i * bits + (uint64_t) k * 64);
      const uint64_t rem = bits - (uint64_t) k * 64;
      if (rem < 64) v &= (((limb_t) 1 << rem) - 1);
   }
   slab[blk * pitch + k] = v;
}

/* dst block i = src block table[i] (unpacking an all-to-all receive buffer into row order) */
__global__ void __launch_bounds__(256)
k_gather_blocks(limb_t *dst, const limb_t *__restrict__ src, const uint32_t *__restrict__ table,
                uint64_t nblocks, uint32_t pitch)
{
   const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= nblocks * pitch) return;
   const uint64_t blk = t / pitch; const uint32_t k = (uint32_t)(t % pitch);
   dst[blk * pitch + k] = src[(uint64_t) table[blk] * pitch + k];
}

/* res += c at limb 0, rippling (one warp; the ripple leaves the first tile only through all-ones
 * limbs).  *carry_out = 1 if the carry leaves limb total-1. */
__global__ void __launch_bounds__(32)
k_add_small(limb_t *res, uint64_t total, uint32_t c, uint32_t *carry_out)
{
   const uint32_t lane = threadIdx.x;
   uint32_t cin = c;
   for (uint64_t k0 = 0; k0 < total && cin; k0 += 32)
   {
      const uint64_t k = k0 + lane;
      limb_t v = (k < total) ? res[k] : ~(limb_t) 0;
      uint32_t g = 0;
      if (k0 == 0 && lane == 0) { const limb_t nv = v + cin; g = (nv < v); v = nv; cin = 0; }
      else if (k0 == 0) cin = 0;
      cin = __shfl_sync(FULL, cin, 0);
      const uint32_t G = __ballot_sync(FULL, g != 0), P = __ballot_sync(FULL, v == ~(limb_t) 0);
      const uint64_t la = mfft_lookahead(G, P, (k0 == 0) ? 0u : 1u);
      if ((la >> lane) & 1u) v += 1;
      if (k < total) res[k] = v;
      cin = (uint32_t)(la >> 32) & 1u;
   }
   if (lane == 0) *carry_out = cin;
}

/* sum of all coefficient windows covering limb k of the result window: coefficient i occupies bits
 * [i*bits, i*bits + NW) of the full result; limb 0 of the window is the limb at bit base_bit (a
 * multiple of 64) -- the window form used by the sharded recombine */
__device__ __forceinline__ mfft_u128 combine_limb_sum(uint64_t k, const limb_t *__restrict__ slab, uint32_t l,
                                                      uint32_t pitch, uint64_t bits, uint64_t ncoef,
                                                      uint64_t base_bit, int small)
{
   const uint64_t lo_bit = base_bit + k * 64, NWb = (uint64_t) l * 64;
   /* coefficients i with i*bits <= lo_bit+63 and i*bits + NW > lo_bit */
   uint64_t imax, imin;
   if (small)
   {  /* every bit offset of the window fits 32 bits: two 32-bit divisions instead of two 64-bit ones */
      const uint32_t lb = (uint32_t) lo_bit, b32 = (uint32_t) bits, nw32 = (uint32_t) NWb;
      imax = (lb + 63u) / b32;
      imin = (lb >= nw32) ? (lb - nw32) / b32 + 1u : 0u;
   } else
   {
      imax = (lo_bit + 63) / bits;
      imin = (lo_bit >= NWb) ? (lo_bit - NWb) / bits + 1 : 0;
   }
   if (imax >= ncoef) imax = ncoef - 1;
   mfft_u128 acc = 0;
   for (uint64_t i = imin; i <= imax && i < ncoef; i++)
   {
      const limb_t *c = slab + i * pitch;
      const uint64_t start = i * bits;            /* bit offset of coefficient i in the result */
      limb_t v;
      if (start <= lo_bit) v = bits_at(c, l, lo_bit - start);
      else v = c[0] << (start - lo_bit);          /* coefficient begins inside this limb */
      acc += v;
   }
   return acc;
}

/* res[k] = low 64 bits of the sum of all coefficient windows covering limb k; cvec[k+1] = the
 * high part of that sum (a small count) */
__global__ void __launch_bounds__(256)
k_combine_sum(limb_t *res, uint32_t *cvec, uint64_t total, const limb_t *__restrict__ slab,
              uint32_t l, uint32_t pitch, uint64_t bits, uint64_t ncoef, uint64_t base_bit, int small)
{
   pdl_wait();
   const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
   if (k > total) return;
   if (k == 0) cvec[0] = 0;
   if (k == total) return;
   const mfft_u128 acc = combine_limb_sum(k, slab, l, pitch, bits, ncoef, base_bit, small);
   res[k] = (limb_t) acc;
   cvec[k + 1] = (uint32_t)(acc >> 64);
}

#define CMB_M 8       /* limbs per lane in the carry passes: one warp tile = 256 limbs */

/* pass 2: res += cvec (tile-local, carry-in 0); per tile: generate bit and propagate bit */
__global__ void __launch_bounds__(128)
k_combine_add(limb_t *res, const uint32_t *__restrict__ cvec, uint64_t total, uint32_t *tileG,
              uint32_t *tileP, uint64_t ntiles)
{
   pdl_wait();
   const uint64_t tile = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const uint32_t lane = threadIdx.x & 31;
   if (tile >= ntiles) return;
   const uint64_t kb = tile * (32 * CMB_M) + (uint64_t) lane * CMB_M;
   limb_t s[CMB_M]; uint32_t c = 0; bool ones_all = true;
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      limb_t x = ~(limb_t) 0, z = 0;
      if (k < total) { x = res[k]; z = cvec[k]; }
      const mfft_u128 acc = (mfft_u128) x + z + c;
      s[i] = (limb_t) acc; c = (uint32_t)(acc >> 64);
      ones_all = ones_all && (s[i] == ~(limb_t) 0);
   }
   const uint32_t G = __ballot_sync(FULL, c != 0), P = __ballot_sync(FULL, ones_all);
   const uint64_t la = mfft_lookahead(G, P, 0);
   uint32_t myc = (uint32_t)(la >> lane) & 1u;
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      const limb_t v = s[i] + myc;
      myc = (myc && v == 0) ? 1u : 0u;
      if (k < total) res[k] = v;
      s[i] = v;
   }
   /* tile generates iff carry out with cin 0; propagates iff every limb is now all ones */
   bool all1 = true;
#pragma unroll
   for (int i = 0; i < CMB_M; i++) all1 = all1 && (kb + i >= total || s[i] == ~(limb_t) 0);
   const bool tp = __all_sync(FULL, all1);
   if (lane == 0) { tileG[tile] = (uint32_t)(la >> 32) & 1u; tileP[tile] = tp ? 1u : 0u; }
}

/* passes 1 + 2 in one kernel (MPIRFFT_COMBINE_FUSED=1): a lane sums the windows of its CMB_M
 * consecutive limbs in registers, takes the high part of the limb before its first one from the
 * previous lane (lane 0 recomputes that one limb) and runs the tile-local carry pass of
 * k_combine_add on the spot -- no cvec array, no intermediate result.  cvec_last[0] receives the high
 * part of limb total-1 (what k_combine_scan reads at cvec[total]). */
__global__ void __launch_bounds__(128)
k_combine_sumadd(limb_t *res, uint32_t *cvec_last, uint64_t total, const limb_t *__restrict__ slab,
                 uint32_t l, uint32_t pitch, uint64_t bits, uint64_t ncoef, uint64_t base_bit, int small,
                 uint32_t *tileG, uint32_t *tileP, uint64_t ntiles)
{
   pdl_wait();
   const uint64_t tile = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const uint32_t lane = threadIdx.x & 31;
   if (tile >= ntiles) return;
   const uint64_t kb = tile * (32 * CMB_M) + (uint64_t) lane * CMB_M;
   limb_t lo[CMB_M]; uint32_t hi[CMB_M];
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      lo[i] = ~(limb_t) 0; hi[i] = 0;
      if (k < total)
      {
         const mfft_u128 acc = combine_limb_sum(k, slab, l, pitch, bits, ncoef, base_bit, small);
         lo[i] = (limb_t) acc; hi[i] = (uint32_t)(acc >> 64);
         if (k + 1 == total) cvec_last[0] = hi[i];
      }
   }
   uint32_t hprev = __shfl_up_sync(FULL, hi[CMB_M - 1], 1);
   if (lane == 0)
      hprev = (kb > 0 && kb - 1 < total) ? (uint32_t)(combine_limb_sum(kb - 1, slab, l, pitch, bits, ncoef, base_bit, small) >> 64) : 0u;
   limb_t s[CMB_M]; uint32_t c = 0; bool ones_all = true;
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      const uint32_t z = (k < total) ? (i ? hi[i - 1] : hprev) : 0u;
      const mfft_u128 acc = (mfft_u128) lo[i] + z + c;
      s[i] = (limb_t) acc; c = (uint32_t)(acc >> 64);
      ones_all = ones_all && (s[i] == ~(limb_t) 0);
   }
   const uint32_t G = __ballot_sync(FULL, c != 0), P = __ballot_sync(FULL, ones_all);
   const uint64_t la = mfft_lookahead(G, P, 0);
   uint32_t myc = (uint32_t)(la >> lane) & 1u;
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      const limb_t v = s[i] + myc;
      myc = (myc && v == 0) ? 1u : 0u;
      if (k < total) res[k] = v;
      s[i] = v;
   }
   bool all1 = true;
#pragma unroll
   for (int i = 0; i < CMB_M; i++) all1 = all1 && (kb + i >= total || s[i] == ~(limb_t) 0);
   const bool tp = __all_sync(FULL, all1);
   if (lane == 0) { tileG[tile] = (uint32_t)(la >> 32) & 1u; tileP[tile] = tp ? 1u : 0u; }
}

/* pass 3: one CTA scans the tile generate/propagate bits.  Each of the 32 warps owns a contiguous
 * group of tiles and walks it 32 tiles per step twice: first with carry-in 0 and 1 to get the
 * group's own generate / propagate, then -- after warp 0 has looked ahead over the 32 groups --
 * with its true carry-in, writing the per-tile carries. */
__global__ void __launch_bounds__(1024)
k_combine_scan(const uint32_t *__restrict__ tileG, const uint32_t *__restrict__ tileP,
               uint32_t *tileC, uint64_t ntiles, const uint32_t *__restrict__ cvec_last, uint32_t *carry_out)
{
   pdl_wait();
   __shared__ uint32_t sG[32], sP[32], sC[33];
   const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
   const uint64_t per = (((ntiles + nwarps - 1) / nwarps + 31) / 32) * 32;     /* tiles per group, multiple of 32 */
   const uint64_t lo = (uint64_t) warp * per, hi = (lo + per < ntiles) ? lo + per : ntiles;
   uint32_t c0 = 0, c1 = 1;
   for (uint64_t t0 = lo; t0 < hi; t0 += 32)
   {
      const uint64_t t = t0 + lane;
      const uint32_t g = (t < hi) ? tileG[t] : 0u, p = (t < hi) ? tileP[t] : 1u;
      const uint32_t G = __ballot_sync(FULL, g != 0), P = __ballot_sync(FULL, p != 0);
      c0 = (uint32_t)(mfft_lookahead(G, P, c0) >> 32) & 1u;
      c1 = (uint32_t)(mfft_lookahead(G, P, c1) >> 32) & 1u;
   }
   if (lane == 0) { sG[warp] = c0; sP[warp] = (c1 && !c0) ? 1u : 0u; }
   __syncthreads();
   if (warp == 0)
   {
      const uint32_t g = (lane < nwarps) ? sG[lane] : 0u, p = (lane < nwarps) ? sP[lane] : 1u;
      const uint32_t G = __ballot_sync(FULL, g != 0), P = __ballot_sync(FULL, p != 0);
      const uint64_t la = mfft_lookahead(G, P, 0);
      sC[lane] = (uint32_t)(la >> lane) & 1u;
      if (lane == 0) sC[32] = (uint32_t)(la >> 32) & 1u;
   }
   __syncthreads();
   uint32_t cin = sC[warp];
   for (uint64_t t0 = lo; t0 < hi; t0 += 32)
   {
      const uint64_t t = t0 + lane;
      const uint32_t g = (t < hi) ? tileG[t] : 0u, p = (t < hi) ? tileP[t] : 1u;
      const uint32_t G = __ballot_sync(FULL, g != 0), P = __ballot_sync(FULL, p != 0);
      const uint64_t la = mfft_lookahead(G, P, cin);
      if (t < hi) tileC[t] = (uint32_t)(la >> lane) & 1u;
      cin = (uint32_t)(la >> 32) & 1u;
   }
   /* what leaves the window: the ripple carry plus the high part of the last limb's column sum */
   if (threadIdx.x == 0 && carry_out) *carry_out = sC[32] + *cvec_last;
}

/* pass 4: tiles with carry-in 1 add it (it ripples through the all-ones prefix of the tile) */
__global__ void __launch_bounds__(128)
k_combine_fix(limb_t *res, uint64_t total, const uint32_t *__restrict__ tileC, uint64_t ntiles)
{
   pdl_wait();
   const uint64_t tile = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const uint32_t lane = threadIdx.x & 31;
   if (tile >= ntiles) return;
   if (!tileC[tile]) return;
   const uint64_t kb = tile * (32 * CMB_M) + (uint64_t) lane * CMB_M;
   limb_t s[CMB_M]; bool ones_all = true;
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      s[i] = (k < total) ? res[k] : ~(limb_t) 0;
      ones_all = ones_all && (s[i] == ~(limb_t) 0);
   }
   const uint32_t P = __ballot_sync(FULL, ones_all);
   const uint64_t la = mfft_lookahead(0, P, 1);
   uint32_t myc = (uint32_t)(la >> lane) & 1u;
#pragma unroll
   for (int i = 0; i < CMB_M; i++)
   {
      const uint64_t k = kb + i;
      const limb_t v = s[i] + myc;
      myc = (myc && v == 0) ? 1u : 0u;
      if (k < total) res[k] = v;
   }
}

/* ------------------------------------------------------------------------------------------ */
/* batched mulmod 2^(64 L) + 1 through a negacyclic transform (FFT_mulmod_2expp1, mul_fft.c:2998) */
/* ------------------------------------------------------------------------------------------ */
/* piece k of operand `prod`: limbs [k*pl, (k+1)*pl) zero-extended to a block (3038-3040, 3049-3051) */
__global__ void __launch_bounds__(256)
k_mm_split(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *__restrict__ src, uint32_t src_pitch,
           uint32_t K, uint32_t pl, uint64_t nblocks)
{
   const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
   const uint64_t per = (uint64_t) l + 1;
   if (t >= nblocks * per) return;
   const uint64_t blk = t / per; const uint32_t k = (uint32_t)(t % per);
   const uint64_t prod = blk / K; const uint32_t piece = (uint32_t)(blk % K);
   slab[blk * pitch + k] = (k < pl) ? src[prod * src_pitch + (uint64_t) piece * pl + k] : 0;
}

/* Recombination of one product per CTA.  C holds the K coefficients of the negacyclic convolution
 * reduced mod 2^(128 pl) + 1 (canonical, l = 2 pl limbs); the true coefficient is
 *      c_k = C_k + m_k (2^(128 pl) + 1),   m_k = (c_k mod 2^64) - C_k[0]   (a small signed number),
 * where c_k mod 2^64 comes from the negacyclic convolution of the pieces' low limbs
 * (fft_naive_convolution_1, 2981-2996; correction 3087-3098).  The result
 *      sum_k c_k B^(pl k)   mod B^L + 1,  L = K pl
 * is built chunk by chunk (chunk q = limbs [q pl, (q+1) pl)): low half of C_q plus high half of
 * C_(q-1) (minus the high half of C_(K-1) for q = 0, B^L = -1), then the small signed terms m_q,
 * m_(q-2) + top_(q-2) and the chunk carries are injected at the chunk's first limb. */
__global__ void __launch_bounds__(256)
k_mm_finish(limb_t *r, uint32_t r_pitch, const limb_t *a, const limb_t *b, uint32_t ab_pitch,
            const limb_t *__restrict__ C, uint32_t l, uint32_t c_pitch, uint32_t K, uint32_t pl)
{
   MFFT_DYN_SMEM(limb_t, sm);
   limb_t *a0 = sm, *b0 = sm + K;
   int64_t *mm = (int64_t *)(sm + 2 * K), *tm = mm + K, *F = tm + K, *F2 = F + K + 1;
   int *flag = (int *)(F2 + K + 1);
   const uint32_t L = K * pl, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
   const uint64_t prod = blockIdx.x;
   const limb_t *A = a + prod * ab_pitch, *Bp = b + prod * ab_pitch;
   limb_t *R = r + prod * r_pitch;
   const limb_t *Cp = C + prod * K * c_pitch;
   const limb_t topA = A[L], topB = Bp[L];
   for (uint32_t k = tid; k < K; k += blockDim.x) { a0[k] = A[(uint64_t) k * pl]; b0[k] = Bp[(uint64_t) k * pl]; }
   __syncthreads();
   if (topA | topB)
   {  /* an operand equal to 2^(64 L) == -1 (its limbs are zero): the product is the negated other
         operand, or 1 (the reference's large path ignores the top limbs, TODO:213-221) */
      const limb_t *X = topA ? Bp : A;
      if (topA && topB) { for (uint32_t k = tid; k <= L; k += blockDim.x) R[k] = (k == 0); return; }
      for (uint32_t k = tid; k < L; k += blockDim.x) R[k] = ~X[k];
      __syncthreads();
      if (tid == 0)
      {  /* -x = ~x + 2 (mod B^L + 1); the ripple leaves limb 0 only through all-ones limbs */
         limb_t c = 2; uint32_t pos = 0;
         while (c && pos < L) { const limb_t v = R[pos], nv = v + c; R[pos] = nv; c = (nv < v); pos++; }
         R[L] = 0;
         if (c)
         {  /* x was 0 or 1: ~x + 2 = B^L or B^L + 1 */
            bool zero = true;
            for (uint32_t k = 0; k < L && zero; k++) zero = (R[k] == 0);
            if (zero) R[L] = 1;      /* x == 1: result B^L */
            /* x == 0 gives B^L + 1 == 0: limbs are (1, 0, ...) with carry -> subtract 1 */
            else { R[0] -= 1; }
         }
      }
      return;
   }
   /* negacyclic convolution of the low limbs mod 2^64 and the correction terms */
   for (uint32_t k = tid; k < K; k += blockDim.x)
   {
      limb_t acc = 0;
      for (uint32_t i = 0; i <= k; i++) acc += a0[i] * b0[k - i];
      for (uint32_t i = k + 1; i < K; i++) acc -= a0[i] * b0[K + k - i];
      const limb_t *Ck = Cp + (uint64_t) k * c_pitch;
      const int64_t m = (int64_t)(acc - Ck[0]);
      mm[k] = m; tm[k] = m + (int64_t) Ck[l];
   }
   __syncthreads();
   /* chunk sums; F[q] = what still has to be added at limb q*pl, F[K] = carry out of the top */
   for (uint32_t q = warp; q < K; q += nwarps)
   {
      const limb_t *X = Cp + (uint64_t) q * c_pitch;
      const limb_t *Y = Cp + (uint64_t)(q ? q - 1 : K - 1) * c_pitch + pl;
      const limb_t mask = q ? 0 : ~(limb_t) 0;
      uint32_t cin = q ? 0u : 1u;
      for (uint32_t k0 = 0; k0 < pl; k0 += 32)
      {
         const limb_t x = X[k0 + lane], y = Y[k0 + lane] ^ mask;
         limb_t sv = x + y;
         const uint32_t G = __ballot_sync(FULL, sv < x), P = __ballot_sync(FULL, sv == ~(limb_t) 0);
         const uint64_t la = mfft_lookahead(G, P, cin);
         sv += (limb_t)((la >> lane) & 1u);
         cin = (uint32_t)(la >> 32) & 1u;
         R[(uint64_t) q * pl + k0 + lane] = sv;
      }
      if (lane == 0)
      {
         const int64_t kq = (int64_t) cin - (q ? 0 : 1);          /* chunk carry, into chunk q+1 */
         const uint32_t q1 = q + 1;
         int64_t e = kq;
         if (q1 < K) e += mm[q1] + (q1 >= 2 ? tm[q1 - 2] : -tm[K - 2 + q1]);
         F[q1] = e;
      }
   }
   if (tid == 0) F[0] = mm[0] - tm[K - 2];          /* chunk 0's own small terms */
   __syncthreads();
   /* inject the small signed terms, rippling inside the chunk; what leaves a chunk goes round again */
   int64_t top = 0;
   for (uint32_t round = 0; round < K + 2; round++)       /* a carry can cross at most K chunks */
   {
      int64_t *Fa = (round & 1) ? F2 : F, *Fb = (round & 1) ? F : F2;
      if (tid == 0) *flag = 0;
      for (uint32_t q = tid; q <= K; q += blockDim.x) Fb[q] = 0;
      __syncthreads();
      for (uint32_t q = tid; q < K; q += blockDim.x)
      {
         int64_t c = Fa[q];
         for (uint32_t pos = 0; c != 0 && pos < pl; pos++)
         {
            const mfft_i128 acc = (mfft_i128) c + (mfft_i128)(mfft_u128) R[(uint64_t) q * pl + pos];
            R[(uint64_t) q * pl + pos] = (limb_t) acc; c = (int64_t)(acc >> 64);
         }
         if (c != 0) { Fb[q + 1] = c; *flag = 1; }
      }
      __syncthreads();
      top += Fa[K];                       /* every thread keeps the same running top */
      const int again = *flag;
      __syncthreads();
      if (!again) { top += Fb[K]; break; }
   }
   /* value = body + top*B^L == body - top: mpn_normmod_2expp1 (272-294) */
   if (tid == 0)
   {
      for (int it = 0; it < 4 && top != 0; it++)
      {
         if (top == 1)
         {
            bool zero = true;
            for (uint32_t k = 0; k < L && zero; k++) zero = (R[k] == 0);
            if (zero) break;
         }
         int64_t c = -top; top = 0;
         for (uint32_t pos = 0; c != 0; pos++)
         {
            if (pos == L) { top = c; break; }
            const mfft_i128 acc = (mfft_i128) c + (mfft_i128)(mfft_u128) R[pos];
            R[pos] = (limb_t) acc; c = (int64_t)(acc >> 64);
         }
      }
      R[L] = (limb_t) top;
   }
}

int mfft_dev_mm_split(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *src, uint32_t src_pitch,
                      uint32_t K, uint32_t pl, uint64_t count, void *stream)
{
   const uint64_t nblocks = count * K, threads = nblocks * ((uint64_t) l + 1);
   if (!threads) return 0;
   PROF(PC_SPLIT, stream);
   MFFT_LAUNCH(k_mm_split, (unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t) stream, slab, l, pitch, src, src_pitch, K, pl, nblocks);
   CKL();
   return 0;
}

int mfft_dev_mm_finish(limb_t *r, uint32_t r_pitch, const limb_t *a, const limb_t *b, uint32_t ab_pitch,
                       const limb_t *C, uint32_t l, uint32_t c_pitch, uint32_t K, uint32_t pl, uint64_t count, void *stream)
{
   if (!count) return 0;
   const size_t sm = ((size_t) 2 * K + 2 * K + 2 * (K + 1)) * 8 + 16;
   PROF(PC_COMBINE, stream);
   CK(cudaFuncSetAttribute(k_mm_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
   MFFT_LAUNCH(k_mm_finish, (unsigned) count, 256, sm, (cudaStream_t) stream, r, r_pitch, a, b, ab_pitch, C, l, c_pitch, K, pl);
   CKL();
   return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* launch ABI                                                                                  */
/* ------------------------------------------------------------------------------------------ */
extern "C" {

#ifdef MFFT_EMU
double mfft_dev_imad_rate(int) { return -1.0; }     /* the microbenchmark lives in imad_peak.cu */
#endif

const char *mfft_dev_last_error(void) { return g_err; }

void mfft_dev_profile_enable(int on) { g_prof_on = on; }
int  mfft_dev_profile_is_on(void) { return g_prof_on; }
void mfft_dev_profile_bytes(double bytes) { if (g_prof_on) g_prof_bytes_next = bytes; }
/* drain the recorded launches: per class total ms, launch count, algorithmic bytes */
int mfft_dev_profile_read(double *ms, uint64_t *launches, double *bytes, int nclass)
{
   for (int i = 0; i < nclass; i++) { ms[i] = 0; launches[i] = 0; bytes[i] = 0; }
#ifndef MFFT_EMU
   if (cudaDeviceSynchronize() != cudaSuccess) return -1;
   for (size_t i = 0; i < g_prof.size(); i++)
   {
      float t = 0; cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b);
      if (g_prof[i].cls < nclass) { ms[g_prof[i].cls] += t; launches[g_prof[i].cls]++; bytes[g_prof[i].cls] += g_prof[i].bytes; }
      cudaEventDestroy(g_prof[i].a); cudaEventDestroy(g_prof[i].b);
   }
   g_prof.clear();
#endif
   return 0;
}
uint64_t mfft_dev_launch_count(void) { return g_launches; }
void mfft_dev_launch_count_reset(void) { g_launches = 0; }

int mfft_dev_count(void)
{
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
   return n;
}

static int g_device = -1;
int mfft_dev_bind(void)
{  /* the current device is per host thread: every entry point rebinds before it allocates or launches */
   if (g_device < 0) { snprintf(g_err, sizeof g_err, "device not initialised"); return -1; }
   CK(cudaSetDevice(g_device));
   return 0;
}

int mfft_dev_init(int device)
{
   int n = mfft_dev_count();
   if (n <= 0) { snprintf(g_err, sizeof g_err, "no CUDA device visible (this library has no CPU path)"); return -1; }
   if (device < 0 || device >= n) { snprintf(g_err, sizeof g_err, "device %d out of range (%d visible)", device, n); return -1; }
   CK(cudaSetDevice(device));
   CK(cudaFree(0));
   g_device = device;
   return 0;
}

void *mfft_dev_alloc(size_t bytes)
{
   void *p = NULL;
   if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess)
   {
      snprintf(g_err, sizeof g_err, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
      return NULL;
   }
   return p;
}
void mfft_dev_free(void *p) { if (p) cudaFree(p); }
void *mfft_host_alloc_pinned(size_t bytes)
{
   void *p = NULL;
   if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return NULL; }
   return p;
}
void mfft_host_free_pinned(void *p) { if (p) cudaFreeHost(p); }

int mfft_dev_h2d(void *d, const void *h, size_t bytes, void *stream)
{ CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream)); return 0; }
int mfft_dev_d2h(void *h, const void *d, size_t bytes, void *stream)
{ CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream)); return 0; }
int mfft_dev_h2d_2d(void *d, size_t dpitch, const void *h, size_t hpitch, size_t width, size_t rows, void *stream)
{ CK(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width, rows, cudaMemcpyHostToDevice, (cudaStream_t) stream)); return 0; }
int mfft_dev_d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width, size_t rows, void *stream)
{ CK(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, rows, cudaMemcpyDeviceToHost, (cudaStream_t) stream)); return 0; }
int mfft_dev_d2d(void *d, const void *s, size_t bytes, void *stream)
{ if (bytes) CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t) stream)); return 0; }
int mfft_dev_memset0(void *d, size_t bytes, void *stream)
{ CK(cudaMemsetAsync(d, 0, bytes, (cudaStream_t) stream)); return 0; }
int mfft_dev_sync(void *stream)
{ CK(cudaStreamSynchronize((cudaStream_t) stream)); return 0; }

/* streams and events for the host-pointer entry points (copies overlapped with the transforms) */
void *mfft_dev_stream_create(void)
{ cudaStream_t s = NULL; if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return NULL; return (void *) s; }
void mfft_dev_stream_destroy(void *s) { if (s) cudaStreamDestroy((cudaStream_t) s); }
void *mfft_dev_event_create(void)
{ cudaEvent_t e = NULL; if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return NULL; return (void *) e; }
void mfft_dev_event_destroy(void *e) { if (e) cudaEventDestroy((cudaEvent_t) e); }
int mfft_dev_event_record(void *e, void *stream) { CK(cudaEventRecord((cudaEvent_t) e, (cudaStream_t) stream)); return 0; }
int mfft_dev_stream_wait(void *stream, void *e) { CK(cudaStreamWaitEvent((cudaStream_t) stream, (cudaEvent_t) e, 0)); return 0; }
int mfft_dev_event_sync(void *e) { return cudaEventSynchronize((cudaEvent_t) e) == cudaSuccess ? 0 : -1; }   /* called from worker threads: no g_err */
int mfft_dev_host_is_pinned(const void *p)
{
   cudaPointerAttributes at;
   if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
   return at.type != cudaMemoryTypeUnregistered;
}

static int pick_m(uint32_t l)
{
   if (l % 256 == 0) return 8;
   if (l % 128 == 0) return 4;
   if (l % 64 == 0) return 2;
   return 1;
}

int mfft_dev_run_stage(limb_t *slab, const mfft_geom *g, const mfft_op *d_ops, uint32_t count,
                       const mfft_batch *d_batch, uint32_t nbatch, void *stream)
{
   if (!count || !nbatch) return 0;
   const uint64_t warps = (uint64_t) count * nbatch;
   const unsigned grid = (unsigned)((warps + 3) / 4);
   cudaStream_t st = (cudaStream_t) stream;
   PROF(PC_STAGE, st);
   switch (pick_m(g->l))
   {
   case 8: MFFT_LAUNCH(k_run_stage<8>, grid, 128, 0, st, slab, *g, d_ops, count, d_batch, nbatch); break;
   case 4: MFFT_LAUNCH(k_run_stage<4>, grid, 128, 0, st, slab, *g, d_ops, count, d_batch, nbatch); break;
   case 2: MFFT_LAUNCH(k_run_stage<2>, grid, 128, 0, st, slab, *g, d_ops, count, d_batch, nbatch); break;
   default: MFFT_LAUNCH(k_run_stage<1>, grid, 128, 0, st, slab, *g, d_ops, count, d_batch, nbatch); break;
   }
   CKL();
   return 0;
}


/* developer aid: per-CTA phase time stamps of the next k_run_tiles launches (globaltimer ns:
 * start, loads issued, loaded, ops done, stored; smid), 8 words per CTA */
static unsigned long long *g_tile_timing = NULL;
extern "C" void mfft_dev_tile_timing(void *buf) { g_tile_timing = (unsigned long long *) buf; }


/* ---- carry-save stage path (mfft_cs_stage.h): any even l; cw = side array of l/2 carry words per block ---- */
int mfft_dev_cs_init(const limb_t *slab, int32_t *cw, const mfft_geom *g, const mfft_batch *d_batch, uint32_t nbatch, void *stream)
{
   const uint64_t total = (uint64_t) g->S * nbatch * (g->l / 2);
   if (!total) return 0;
   MFFT_LAUNCH(k_cs_init, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t) stream, slab, cw, *g, d_batch, nbatch);
   CKL();
   return 0;
}

int mfft_dev_run_stage_cs(limb_t *slab, int32_t *cw, const mfft_geom *g, const mfft_op *d_ops, uint32_t count,
                          const mfft_batch *d_batch, uint32_t nbatch, void *stream)
{
   if (!count || !nbatch) return 0;
   PROF(PC_STAGE, stream);
   /* 64 / 128 / 256 threads per CTA measure the same: the layer is HBM-bound (about 5.8 TB/s of slab +
      carry-word traffic at 2^26 limbs) */
   MFFT_LAUNCH(k_stage_cs, (unsigned)((uint64_t) count * nbatch), 256, 0, (cudaStream_t) stream, slab, cw, *g, d_ops, count, d_batch, nbatch);
   CKL();
   return 0;
}

int mfft_dev_finalize_cs(limb_t *dst, uint32_t dst_stride, const uint32_t *d_dst_base, limb_t *slab, int32_t *cw,
                         const mfft_geom *g, const mfft_move *d_moves, uint32_t nmoves, const mfft_batch *d_batch,
                         uint32_t nbatch, int normalise, void *stream)
{
   if (!nmoves || !nbatch) return 0;
   const size_t sm = ((size_t) g->l + 15) & ~(size_t) 15;          /* 2 bytes per chunk */
   PROF(PC_FINALIZE, stream);
   CK(cudaFuncSetAttribute(k_finalize_cs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
   MFFT_LAUNCH(k_finalize_cs, (unsigned)((uint64_t) nmoves * nbatch), 256, sm, (cudaStream_t) stream, dst, dst_stride, d_dst_base,
               slab, cw, *g, d_moves, nmoves, d_batch, nbatch, normalise);
   CKL();
   return 0;
}

int mfft_dev_run_stage_cs_ip(limb_t *slab, int32_t *cw, const mfft_geom *g, const mfft_op *d_ops, uint32_t count,
                             const mfft_batch *d_batch, uint32_t nbatch, uint32_t nstaged, void *stream)
{
   if (!count || !nbatch) return 0;
   /* nstaged: how many operands an op of this stage has to stage at most (0..2) */
   const size_t sm = (size_t)(nstaged > 2 ? 2 : nstaged) * ((size_t) g->l * 8 + (size_t) g->l * 2);
   if (sm > 227 * 1024) { snprintf(g_err, sizeof g_err, "run_stage_cs_ip: l=%u too large for in-place staging", g->l); return -2; }
   /* big coefficients: few CTAs fit an SM, so each gets more threads to keep the loads in flight */
   const unsigned threads = (sm > 96 * 1024) ? 1024u : (sm > 40 * 1024) ? 512u : 256u;
   PROF(PC_STAGE, stream);
   CK(cudaFuncSetAttribute(k_stage_cs_ip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
   MFFT_LAUNCH(k_stage_cs_ip, (unsigned)((uint64_t) count * nbatch), threads, sm, (cudaStream_t) stream, slab, cw, *g, d_ops, count, d_batch, nbatch);
   CKL();
   return 0;
}

size_t mfft_dev_sliced_coeff_bytes(uint32_t nchv)
{
   switch (nchv) { case 32: case 64: case 96: case 128: case 192: case 256: return (size_t) nchv * 16 + 16 + (size_t) nchv * 4; default: return 0; }
}

int mfft_dev_run_tiles_sliced(limb_t *slab, int32_t *cw, const mfft_geom *g, uint32_t gs, uint32_t nchv, uint32_t R,
                              const mfft_tile *d_tiles, uint32_t ntiles, const uint32_t *d_pos, const mfft_tileop *d_ops,
                              const uint32_t *d_stoff, uint32_t max_npos, uint32_t max_nops,
                              const mfft_batch *d_batch, uint32_t nbatch, void *stream)
{
   if (!ntiles || !nbatch || !gs) return 0;
   if (!R || gs % R) { snprintf(g_err, sizeof g_err, "run_tiles_sliced: %u slices per CTA do not divide the stride %u", R, gs); return -2; }
   const size_t cb = mfft_dev_sliced_coeff_bytes(nchv);
   if (!cb) { snprintf(g_err, sizeof g_err, "run_tiles_sliced: %u chunks per slice unsupported", nchv); return -2; }
   const uint32_t desc = (uint32_t)(((size_t) max_nops * sizeof(mfft_tileop) + (size_t) max_npos * 4 + 4 * 64 + 15) & ~(size_t) 15);
   const size_t smem = desc + (size_t) max_npos * cb * R;
   const uint64_t grid = (uint64_t) ntiles * nbatch * (gs / R);
   cudaStream_t st = (cudaStream_t) stream;
   if (smem > 227 * 1024 || grid > 0x7fffffffull) { snprintf(g_err, sizeof g_err, "run_tiles_sliced: launch too large"); return -2; }
   PROF(PC_STAGE, st);
#define RUN_SLICED(NN, TH)                                                                          \
   do {                                                                                            \
      CK(cudaFuncSetAttribute(k_run_tiles_sliced<NN, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
      MFFT_LAUNCH((k_run_tiles_sliced<NN, TH>), (unsigned) grid, TH, smem, st, slab, cw, *g, gs, R, d_tiles, d_pos, d_ops, d_stoff, \
                  d_batch, nbatch, desc);                                                          \
   } while (0)
   switch (nchv / 32)
   {
   case 1: RUN_SLICED(1, 256); break;
   case 2: RUN_SLICED(2, 256); break;
   case 3: RUN_SLICED(3, 256); break;
   case 4: RUN_SLICED(4, 256); break;
   case 6: RUN_SLICED(6, 256); break;
   default: RUN_SLICED(8, 256); break;
   }
#undef RUN_SLICED
   CKL();
   return 0;
}

/* coefficient sizes the fused executor is instantiated for: l = 64*NT limbs */
static int tiles_cfg(uint32_t l, int *NT)
{
   switch (l)
   {
   case 64: case 128: case 192: case 256: case 384: case 512: *NT = (int)(l / 64); return 1;
   default: return 0;
   }
}

int mfft_dev_tiles_supported(uint32_t l) { int nt; return tiles_cfg(l, &nt); }

static size_t tiles_coeff_bytes(uint32_t l)
{
   /* the block image (l body limbs, top limb, pad limb: one 16-byte aligned slab pitch) + l/2 chunks
      with a 32-bit carry word each (mfft_tiles.h: tile_cfg) */
   return (size_t) l * 8 + 16 + (size_t)(l / 2) * 4;
}

uint32_t mfft_dev_tiles_max_npos(uint32_t l)
{
   if (!mfft_dev_tiles_supported(l)) return 0;
   /* Tiles of about 45 KB, at least 16 coefficients (4 radix-2 layers per pass), at most 100 KB:
      several small CTAs per SM overlap their load / compute / store phases better than two large
      ones (measured at l = 256: 16 coefficients per tile, four 4-warp CTAs per SM, same pass count) */
   const size_t cb = tiles_coeff_bytes(l), nmax = (100 * 1024) / cb, npref = (45 * 1024) / cb;
   uint32_t p = 4;
   const char *env = getenv("MPIRFFT_TILE_NPOS");      /* developer aid: force the tile size */
   while (p * 2 <= npref && p * 2 <= 128) p *= 2;
   while (p < 16 && p * 2 <= nmax) p *= 2;
   if (env && atoi(env) >= 4 && (size_t) atoi(env) <= nmax) p = (uint32_t) atoi(env);
   return p;
}

int mfft_dev_run_tiles(limb_t *slab, const mfft_geom *g, const mfft_tile *d_tiles, uint32_t ntiles,
                       const uint32_t *d_pos, const mfft_tileop *d_ops, uint32_t max_npos, uint32_t max_nops,
                       const mfft_batch *d_batch, uint32_t nbatch,
                       limb_t *dst, const uint32_t *d_dstpos, const uint32_t *d_dst_base,
                       uint32_t dst_stride, int normalise, const uint32_t *d_stoff, int heavy,
                        const mfft_tile *h_tiles, const uint32_t *h_pos, const uint32_t *h_stoff,
                        const mfft_batch *h_batch, const mfft_split *split, void *stream)
{
   int NT = 0;
   if (!ntiles || !nbatch) return 0;
   if (!tiles_cfg(g->l, &NT)) { snprintf(g_err, sizeof g_err, "run_tiles: l=%u unsupported", g->l); return -2; }
   /* descriptor area (ops + position list), rounded to 16 bytes, then coefficients */
   const uint32_t desc = (uint32_t)((MBAR_BYTES + (size_t) max_nops * sizeof(mfft_tileop) + (size_t) max_npos * 4 + 4 * 64 + 15) & ~(size_t) 15);
   const size_t smem = desc + (size_t) max_npos * tiles_coeff_bytes(g->l);
   const unsigned grid = ntiles * nbatch;
   cudaStream_t st = (cudaStream_t) stream;
   /* small passes: tile descriptors, position lists and stage offsets travel as kernel parameters */
   tile_params tp;                                      /* per call: launches from several host threads do not share it */
   tp.valid = 0; tp.batch_valid = 0;
   { static int dbg = -1; if (dbg < 0) { const char *e = getenv("MPIRFFT_TILE_DEBUG"); dbg = e ? atoi(e) : 0; } tp.debug = (uint32_t) dbg; }
   tp.split = split ? 1u : 0u; tp.split_src = split ? split->src : NULL; tp.split_nlimbs = split ? split->nlimbs : 0;
   tp.split_bits = split ? split->bits : 0; tp.split_ncoef = split ? split->ncoef : 0;
   if (h_tiles && h_pos && h_stoff && ntiles <= TP_MAXT)
   {
      uint32_t t, ok = 1;
      for (t = 0; t < ntiles; t++) if (h_tiles[t].npos > TP_MAXP || h_tiles[t].nstages + 1 > TP_MAXS) ok = 0;
      if (ok)
      {
         for (t = 0; t < ntiles; t++)
         {
            tp.tiles[t] = h_tiles[t];
            memcpy(tp.pos + t * TP_MAXP, h_pos + h_tiles[t].pos_off, sizeof(uint32_t) * h_tiles[t].npos);
            memcpy(tp.stoff + t * TP_MAXS, h_stoff + h_tiles[t].pad, sizeof(uint32_t) * (h_tiles[t].nstages + 1));
         }
         tp.valid = 1;
      }
   }
   if (h_batch && nbatch <= TP_MAXB) { memcpy(tp.batch, h_batch, sizeof(mfft_batch) * nbatch); tp.batch_valid = 1; }
   const int pdl = pdl_on();
   PROF(PC_STAGE, st);
   /* two CTAs per SM: 16 warps x 64 registers while a lane holds <= 4 chunk pairs of an op, else
      8 warps x 128 registers (fewer, larger coefficients per tile) */
#define RUN_TILES(NN, TH, WS)                                                                      \
   do {                                                                                            \
      CK(cudaFuncSetAttribute(k_run_tiles<NN, TH, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
      MFFT_LAUNCH_PDL(pdl, (k_run_tiles<NN, TH, WS>), grid, TH, smem, st, slab, *g, d_tiles, d_pos, d_ops, d_batch, nbatch, \
                  dst, d_dstpos, d_dst_base, dst_stride, normalise, desc, d_stoff, g_tile_timing, tp); \
   } while (0)
   static int wsplit = -1;       /* MPIRFFT_TILE_WSPLIT=0: one warp per op at l = 256 (the 4-warp CTAs of round 1) */
   if (wsplit < 0) { const char *e = getenv("MPIRFFT_TILE_WSPLIT"); wsplit = e ? atoi(e) : 1; }
   {  /* persistent pipelined variant (k_run_tiles_p): one CTA per SM, a ring of tile buffers filled by
         bulk copies while the warp groups work.  Needs at least two tiles per SM to have anything to
         overlap, four-warp tiles (NT <= 4), and room for the ring.  MPIRFFT_TILE_PERSIST=1 turns it on. */
      static int persist = -1, nsm = 0;
      if (persist < 0)
      {
         const char *e = getenv("MPIRFFT_TILE_PERSIST"); persist = e ? atoi(e) : 0;     /* opt-in: measured slower at 2^20 limbs (7 tiles per SM are too few to pipeline) */
#ifndef MFFT_EMU
         int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
#else
         nsm = 2;
#endif
         if (nsm <= 0) nsm = 148;
      }
      const size_t buf = desc + (size_t) max_npos * tiles_coeff_bytes(g->l);
      const size_t smem_p = TPP_CTL_BYTES + TPP_NBUF * buf;
      const int emu_force = (persist == 2);          /* tests: also for small launches */
      if (persist && NT <= 4 && smem_p <= 227 * 1024 && (grid >= 2u * (unsigned) nsm || emu_force) &&
          (((g->pitch & 1u) == 0) || emu_force))
      {
         const unsigned gp = grid < (unsigned) nsm ? grid : (unsigned) nsm;
#define RUN_TILES_P(NN)                                                                            \
         do {                                                                                      \
            CK(cudaFuncSetAttribute(k_run_tiles_p<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_p)); \
            MFFT_LAUNCH_PDL(pdl, (k_run_tiles_p<NN>), gp, TPP_GROUPS * 128, smem_p, st, slab, *g, d_tiles, d_pos, d_ops, d_batch, nbatch, \
                        dst, d_dstpos, d_dst_base, dst_stride, normalise, desc, d_stoff, grid, (uint32_t) buf, tp); \
         } while (0)
         switch (NT)
         {
         case 1: RUN_TILES_P(1); break;
         case 2: RUN_TILES_P(2); break;
         case 3: RUN_TILES_P(3); break;
         default: RUN_TILES_P(4); break;
         }
#undef RUN_TILES_P
         CKL();
         return 0;
      }
   }
   if (max_npos <= 16 && NT <= 4 && smem <= 56 * 1024)
      switch (NT)          /* small tiles: four CTAs per SM */
      {
      case 1: RUN_TILES(1, 128, 1); break;
      case 2: RUN_TILES(2, 128, 1); break;
      case 3: RUN_TILES(3, 128, 1); break;
      default: if (wsplit) RUN_TILES(4, 256, 2); else RUN_TILES(4, 128, 1); break;
      }
   else if (heavy && NT <= 4)
      switch (NT)
      {
      case 1: RUN_TILES(1, 256, 1); break;
      case 2: RUN_TILES(2, 256, 1); break;
      case 3: RUN_TILES(3, 256, 1); break;
      default: RUN_TILES(4, 256, 1); break;
      }
   else
   switch (NT)
   {
   case 1: RUN_TILES(1, 512, 1); break;
   case 2: RUN_TILES(2, 512, 1); break;
   case 3: RUN_TILES(3, 512, 1); break;
   case 4: RUN_TILES(4, 512, 1); break;
   case 6: RUN_TILES(6, 256, 1); break;
   default: RUN_TILES(8, 256, 1); break;
   }
#undef RUN_TILES
   if (g_tile_timing) g_tile_timing += (size_t) grid * 8;       /* next launch stamps behind this one */
   CKL();
   return 0;
}

int mfft_dev_finalize(limb_t *dst, uint32_t dst_stride, const uint32_t *d_dst_base,
                      const limb_t *slab, const mfft_geom *g, const mfft_move *d_moves, uint32_t nmoves,
                      const mfft_batch *d_batch, uint32_t nbatch, uint32_t shift, int normalise,
                      void *stream)
{
   if (!nmoves || !nbatch) return 0;
   const uint64_t warps = (uint64_t) nmoves * nbatch;
   const unsigned grid = (unsigned)((warps + 3) / 4);
   cudaStream_t st = (cudaStream_t) stream;
   limb_t *s = (limb_t *) slab;
   PROF(PC_FINALIZE, st);
   switch (pick_m(g->l))
   {
   case 8: MFFT_LAUNCH(k_finalize<8>, grid, 128, 0, st, dst, dst_stride, d_dst_base, s, *g, d_moves, nmoves, d_batch, nbatch, shift, normalise); break;
   case 4: MFFT_LAUNCH(k_finalize<4>, grid, 128, 0, st, dst, dst_stride, d_dst_base, s, *g, d_moves, nmoves, d_batch, nbatch, shift, normalise); break;
   case 2: MFFT_LAUNCH(k_finalize<2>, grid, 128, 0, st, dst, dst_stride, d_dst_base, s, *g, d_moves, nmoves, d_batch, nbatch, shift, normalise); break;
   default: MFFT_LAUNCH(k_finalize<1>, grid, 128, 0, st, dst, dst_stride, d_dst_base, s, *g, d_moves, nmoves, d_batch, nbatch, shift, normalise); break;
   }
   CKL();
   return 0;
}

int mfft_dev_normalise(limb_t *slab, uint32_t l, uint32_t pitch, uint64_t nblk, void *stream)
{
   if (!nblk) return 0;
   PROF(PC_NORMALISE, stream);
   MFFT_LAUNCH(k_normalise, (unsigned)((nblk + 3) / 4), 128, 0, (cudaStream_t) stream, slab, l, pitch, nblk);
   CKL();
   return 0;
}

static limb_t *g_pw_scratch = NULL; static size_t g_pw_scratch_bytes = 0;
static int g_pw_mode = -1;     /* 0 auto, 1 nested SS, 2 Karatsuba-split blocks, 3 schoolbook blocks, 4 / 5 two Karatsuba levels */
#define PW_KARA_DEFAULT(l) ((l) == 512 || (l) == 256 || (l) == 128)
void mfft_dev_pointwise_mode(int mode) { g_pw_mode = mode; }

/* warps of a product kernel (launched as four-warp CTAs with `smem` bytes each) that are resident on the
   device at once; a multiple of 4 */
static unsigned pw_capacity(const void *kern, size_t smem)
{
   static const void *fk[16]; static unsigned fv[16]; static int nf = 0;
   for (int i = 0; i < nf; i++) if (fk[i] == kern) return fv[i];
   unsigned cap = 0;
#ifdef MFFT_EMU
   (void) smem;
   { const char *e = getenv("MPIRFFT_PW_TAIL_TEST"); cap = e ? 8u : 0u; }      /* tests: two "SMs" of one CTA */
#else
   int nb = 0, dev = 0, nsm = 0;
   if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, smem) != cudaSuccess) { cudaGetLastError(); nb = 0; }
   cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
   if (nb > 0 && nsm > 0) cap = (unsigned) nb * 4u * (unsigned) nsm;
#endif
   if (nf < 16) { fk[nf] = kern; fv[nf] = cap; nf++; }
   return cap;
}

int mfft_dev_pointwise(limb_t *a, const limb_t *b, const uint32_t *d_blocks, uint32_t nblk,
                       uint32_t l, uint32_t pitch, void *stream)
{
   if (!nblk) return 0;
   cudaStream_t st = (cudaStream_t) stream;
   const unsigned grid = (nblk + 3) / 4;
   PROF(PC_POINTWISE, st);
   {  /* mode 0 (default): schoolbook IMAD.WIDE kernel; mode 1 (MPIRFFT_POINTWISE=ss or
         mfft_dev_pointwise_mode): the nested Schoenhage-Strassen step inside a warp (k_mulmod_ss).
         Measured at l = 256 on B200: 0.59 ms vs 0.88 ms per 16 640 products, so direct stays default.
         (A reduced-radix variant -- 27-bit digits, carry-less 64-bit column sums -- was tried and is
         slower, 0.91 ms: ptxas never emits IMAD.WIDE with a non-zero 64-bit addend on sm_100a, it
         splits every mad.wide.u32 into IMAD.WIDE(.., RZ) + IADD3 + IADD3.X, so the real ceiling for
         32x32->64 multiply-ADDs is the ~31/clk/SM of the IMAD.WIDE.U32.X chains used here.) */
      if (g_pw_mode < 0) { const char *e = getenv("MPIRFFT_POINTWISE"); g_pw_mode = !e ? 0 : e[0] == 's' ? 1 : e[0] == 'k' ? 2 : e[0] == 'd' ? 3 : e[0] == '2' ? 4 : e[0] == 'm' ? 5 : 0; }
      uint32_t np = 0, lp = 0;
      if (g_pw_mode == 1)
      {
         if (l == 64)  { np = 16; lp = 5; }       /* 32 pieces of 128 bits, ring 2^320+1,  w' = 20 */
         if (l == 128) { np = 32; lp = 5; }       /* 64 pieces of 128 bits, ring 2^320+1,  w' = 10 */
         if (l == 256) { np = 32; lp = 9; }       /* 64 pieces of 256 bits, ring 2^576+1,  w' = 18 */
         if (l == 512) { np = 64; lp = 10; }      /* 128 pieces of 256 bits, ring 2^640+1, w' = 10 */
      }
      if (np)
      {
         const uint32_t wp = 64 * lp / np, cp = (lp + 1) | 1;
         uint32_t depthp = 0; while ((1u << depthp) < np) depthp++;
         const size_t sm = (size_t) 4 * 2 * (2 * np) * cp * 8;
         if (lp == 5) { CK(cudaFuncSetAttribute(k_mulmod_ss<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
                        MFFT_LAUNCH(k_mulmod_ss<5>, grid, 128, sm, st, a, b, d_blocks, nblk, l, pitch, np, wp, depthp); }
         else if (lp == 9) { CK(cudaFuncSetAttribute(k_mulmod_ss<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
                        MFFT_LAUNCH(k_mulmod_ss<9>, grid, 128, sm, st, a, b, d_blocks, nblk, l, pitch, np, wp, depthp); }
         else { CK(cudaFuncSetAttribute(k_mulmod_ss<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
                        MFFT_LAUNCH(k_mulmod_ss<10>, grid, 128, sm, st, a, b, d_blocks, nblk, l, pitch, np, wp, depthp); }
         CKL();
         return 0;
      }
   }
   static int unr = -1;        /* step-loop unroll factor (tuning aid): MPIRFFT_PW_UNROLL = 1 or 4; default 4 */
   if (unr < 0) { const char *e = getenv("MPIRFFT_PW_UNROLL"); unr = e ? atoi(e) : 0; }
   const bool kara = (g_pw_mode == 2) || ((g_pw_mode == 0 || g_pw_mode >= 4) && PW_KARA_DEFAULT(l));
   const int u = (unr == 0 && l == 512 && !kara) ? 1 : unr;   /* schoolbook blocks at l = 512: 252 registers already, unrolling only adds spills */
#define PW_SMEM(CC, KA) ((size_t) 4 * 32 * ((KA) == 2 ? (CC) + (CC) / 2 + 1 + 3 * ((CC) / 4 + 1) + 2 * (CC) + 3 \
                                              : (KA) ? (CC) + (CC) / 2 + 1 + ((CC) >= 32 ? 2 * (CC) + 3 : 0) : (CC)) * 4)
   /* Experiment kept as an opt-in (MPIRFFT_PW_TAIL=1): the remainder of the last, partly filled wave in a
      second launch of one-warp CTAs, which spread over the SMs and have an IMAD pipe each.  Measured at cfg2
      (16 640 products = 14.05 waves): SLOWER, 0.528 -> 0.546 ms -- the partly filled wave already overlaps the
      start of the next pass, and the second launch puts one product's full latency behind the first. */
   static int tailsplit = -1;
   if (tailsplit < 0) { const char *e = getenv("MPIRFFT_PW_TAIL"); tailsplit = e ? atoi(e) : 0; }
#define PW_LAUNCH_K(KERN, SM4) do { \
      const size_t sm__ = (SM4); \
      const unsigned cap__ = tailsplit ? pw_capacity((const void *) KERN, sm__) : 0u; \
      const uint32_t tail__ = (cap__ && nblk > cap__) ? nblk % cap__ : 0u; \
      if (tail__ && tail__ <= cap__ / 4) \
      { \
         MFFT_LAUNCH_PDL(pdl_on(), (KERN), (nblk - tail__) / 4, 128, sm__, st, a, b, d_blocks, nblk - tail__, l, pitch); \
         MFFT_LAUNCH_PDL(pdl_on(), (KERN), tail__, 32, sm__ / 4, st, a, b, d_blocks + (nblk - tail__), tail__, l, pitch); \
         g_launches++;        /* CKL() below counts one launch per call */ \
      } \
      else MFFT_LAUNCH_PDL(pdl_on(), (KERN), grid, 128, sm__, st, a, b, d_blocks, nblk, l, pitch); } while (0)
#define PW_LAUNCH(CC, KA) do { \
      if (u == 1) PW_LAUNCH_K((k_pointwise<CC, KA, 1>), PW_SMEM(CC, KA)); \
      else PW_LAUNCH_K((k_pointwise<CC, KA, 4>), PW_SMEM(CC, KA)); } while (0)
#define PW_LAUNCH2(CC, MG) do { \
      if (u == 1) PW_LAUNCH_K((k_pointwise<CC, 2, 1, MG>), PW_SMEM(CC, 2)); \
      else PW_LAUNCH_K((k_pointwise<CC, 2, 4, MG>), PW_SMEM(CC, 2)); } while (0)
   static bool attr = false;
   if (!attr)
   {  /* the kernels with parked words need more than the default 48 KB per CTA */
      CK(cudaFuncSetAttribute(k_pointwise<32, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) PW_SMEM(32, 1)));
      CK(cudaFuncSetAttribute(k_pointwise<32, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) PW_SMEM(32, 1)));
      CK(cudaFuncSetAttribute(k_pointwise<32, 2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) PW_SMEM(32, 2)));
      CK(cudaFuncSetAttribute(k_pointwise<32, 2, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) PW_SMEM(32, 2)));
      attr = true;
   }
   const int kara2 = (g_pw_mode == 4) ? 1 : (g_pw_mode == 5) ? 2 : 0;     /* two levels; 2: L and Hh in one sweep */
   if (kara2 && l == 256) { if (kara2 == 2) PW_LAUNCH2(16, true); else PW_LAUNCH2(16, false); }
   else if (kara2 && l == 512) PW_LAUNCH2(32, false);
   else if (l == 64)  { if (kara) PW_LAUNCH(4, 1); else PW_LAUNCH(4, 0); }
   else if (l == 128) { if (kara) PW_LAUNCH(8, 1); else PW_LAUNCH(8, 0); }
   else if (l == 256) { if (kara) PW_LAUNCH(16, 1); else PW_LAUNCH(16, 0); }
   else if (l == 512) { if (kara) PW_LAUNCH(32, 1); else PW_LAUNCH(32, 0); }
   else
   {
      const size_t need = (size_t) nblk * 3 * l * sizeof(limb_t);
      const size_t sm = 2 * (size_t) l * sizeof(limb_t);
      if (sm > 200 * 1024) { snprintf(g_err, sizeof g_err, "pointwise: l=%u too large for the direct kernel", l); return -2; }
      if (need > g_pw_scratch_bytes)
      {
         CK(cudaStreamSynchronize(st));
         if (g_pw_scratch) cudaFree(g_pw_scratch);
         g_pw_scratch = NULL; g_pw_scratch_bytes = 0;
         CK(cudaMalloc(&g_pw_scratch, need)); g_pw_scratch_bytes = need;
      }
      CK(cudaFuncSetAttribute(k_pointwise_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
      MFFT_LAUNCH(k_pointwise_generic, nblk, 256, sm, st, a, b, d_blocks, nblk, l, pitch, g_pw_scratch);
   }
   CKL();
   return 0;
}

int mfft_dev_split(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *src, uint64_t nlimbs,
                   uint64_t bits, uint64_t ncoef, uint64_t nzero, void *stream)
{
   if (nzero < ncoef) nzero = ncoef;
   const uint64_t threads = nzero * ((uint64_t) l + 1);
   if (!threads) return 0;
   PROF(PC_SPLIT, stream);
   MFFT_LAUNCH(k_split, (unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t) stream, slab, l, pitch, src, nlimbs, bits, ncoef, nzero, 0u, 0u, 0u);
   CKL();
   return 0;
}

int mfft_dev_split_cols(limb_t *slab, uint32_t l, uint32_t pitch, const limb_t *src, uint64_t nlimbs,
                        uint64_t bits, uint64_t ncoef, uint64_t nrows, uint32_t ncl, uint32_t n1, uint32_t c0, void *stream)
{
   const uint64_t nblocks = nrows * ncl, threads = nblocks * ((uint64_t) l + 1);
   if (!threads) return 0;
   PROF(PC_SPLIT, stream);
   MFFT_LAUNCH(k_split, (unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t) stream, slab, l, pitch, src, nlimbs, bits, ncoef, nblocks, ncl, n1, c0);
   CKL();
   return 0;
}

int mfft_dev_gather_blocks(limb_t *dst, const limb_t *src, const uint32_t *d_table, uint64_t nblocks, uint32_t pitch, void *stream)
{
   const uint64_t threads = nblocks * pitch;
   if (!threads) return 0;
   MFFT_LAUNCH(k_gather_blocks, (unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t) stream, dst, src, d_table, nblocks, pitch);
   CKL();
   return 0;
}

int mfft_dev_add_small(limb_t *res, uint64_t total, uint32_t c, uint32_t *d_carry_out, void *stream)
{
   MFFT_LAUNCH(k_add_small, 1, 32, 0, (cudaStream_t) stream, res, total, c, d_carry_out);
   CKL();
   return 0;
}

size_t mfft_dev_combine_work(uint64_t total)
{
   const uint64_t ntiles = (total + 32 * CMB_M - 1) / (32 * CMB_M);
   return (size_t)((total + 1) * 4 + 3 * (ntiles + 32) * 4 + 256);
}

int mfft_dev_combine_window(limb_t *res, uint64_t total, const limb_t *slab, uint32_t l, uint32_t pitch,
                            uint64_t bits, uint64_t ncoef, uint64_t base_bit, uint32_t *d_carry_out, void *work, void *stream);

int mfft_dev_combine(limb_t *res, uint64_t total, const limb_t *slab, uint32_t l, uint32_t pitch,
                     uint64_t bits, uint64_t ncoef, void *work, void *stream)
{
   return mfft_dev_combine_window(res, total, slab, l, pitch, bits, ncoef, 0, NULL, work, stream);
}

/* res[k], k < total: limb k of the window that starts at bit base_bit of sum_i block_i * 2^(i*bits)
 * (each limb sums the pieces of the coefficient windows that cover it; the carries run inside
 * the window only).  *d_carry_out (optional): the carry that leaves limb total-1. */
int mfft_dev_combine_window(limb_t *res, uint64_t total, const limb_t *slab, uint32_t l, uint32_t pitch,
                            uint64_t bits, uint64_t ncoef, uint64_t base_bit, uint32_t *d_carry_out, void *work, void *stream)
{
   cudaStream_t st = (cudaStream_t) stream;
   if (!total) return 0;
   if (!ncoef) { CK(cudaMemsetAsync(res, 0, total * 8, st)); if (d_carry_out) CK(cudaMemsetAsync(d_carry_out, 0, 4, st)); return 0; }
   const uint64_t ntiles = (total + 32 * CMB_M - 1) / (32 * CMB_M);
   uint32_t *cvec = (uint32_t *) work;
   uint32_t *tileG = cvec + ((total + 1 + 63) / 64) * 64;
   uint32_t *tileP = tileG + ntiles + 32 - (ntiles % 32);
   uint32_t *tileC = tileP + ntiles + 32 - (ntiles % 32);
   PROF(PC_COMBINE, st);
   static int fused = -1;
   if (fused < 0) { const char *e = getenv("MPIRFFT_COMBINE_FUSED"); fused = e ? (e[0] != '0') : MFFT_COMBINE_FUSED_DEFAULT; }
   const int small = (int)(base_bit + (total + 2) * 64 < 0xffffff00ull && bits < 0xffffffffull);
   if (fused)
   {
      MFFT_LAUNCH_PDL(pdl_on(), k_combine_sumadd, (unsigned)((ntiles + 3) / 4), 128, 0, st, res, cvec + total, total, slab, l, pitch,
                      bits, ncoef, base_bit, small, tileG, tileP, ntiles);
      CKL();
   } else
   {
      MFFT_LAUNCH_PDL(pdl_on(), k_combine_sum, (unsigned)((total + 1 + 255) / 256), 256, 0, st, res, cvec, total, slab, l, pitch, bits, ncoef, base_bit, small);
      CKL();
      MFFT_LAUNCH_PDL(pdl_on(), k_combine_add, (unsigned)((ntiles + 3) / 4), 128, 0, st, res, cvec, total, tileG, tileP, ntiles);
      CKL();
   }
   MFFT_LAUNCH_PDL(pdl_on(), k_combine_scan, 1, 1024, 0, st, tileG, tileP, tileC, ntiles, cvec + total, d_carry_out);
   CKL();
   MFFT_LAUNCH_PDL(pdl_on(), k_combine_fix, (unsigned)((ntiles + 3) / 4), 128, 0, st, res, total, tileC, ntiles);
   CKL();
   return 0;
}

} /* extern "C" */
