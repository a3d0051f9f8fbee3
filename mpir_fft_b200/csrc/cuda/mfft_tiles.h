/* mfft_tiles.h -- the fused shared-memory tile executor (included by mfft_kernels.cu).
 *
 * Several radix-2 layers of a transform (mul_fft.c:786-827, 1397-1442, 1128-1227, 1444-1536, ...)
 * run on a tile of coefficients that stays in shared memory; HBM sees one read and one write of
 * the tile per pass.
 *
 * Carry-save coefficients.  Inside a tile a coefficient of l = 64*NT limbs is NCH = 32*NT
 * *chunks* of 2 limbs (128 bits), each with a signed 32-bit *carry word*:
 *
 *        value = sum_j ( x_j + c_j * B^2 ) * B^(2j)      (mod p = B^l + 1,  B = 2^64)
 *
 * i.e. c_j is a deferred carry into chunk j+1, and c_{NCH-1} is the reference's signed top limb
 * (README:54).  Additions never ripple past a chunk: the carry out of the 2-limb add chain is
 * added to the chunk's carry word.  A rotation by a whole number of chunks (every inner MFA
 * layer: the roots are 2^(w*n1*j), 2^(w*n2*j), multiples of 128 bits from 2^20 limbs up) moves
 * chunks together with their carry words, and the negation of a chunk that wraps around p is
 * local too:  -(x + c*B^2) = ~x + 1 + (-1-c)*B^2.  So an aligned butterfly is purely lane-local
 * work -- no ballots, no shuffles, no serial ripple: per lane and chunk two 16-byte loads, two
 * 2-limb add chains, two 16-byte stores.  The carry words are resolved once per pass, when a
 * coefficient goes back to HBM in the reference's block layout.
 *
 * Layout of one coefficient in shared memory: l body limbs (dense, so lane j's chunk is the 16
 * bytes at 16 j: conflict-free LDS.128/STS.128), then the NCH carry words.
 *
 * Rotations that are not chunk-aligned (the z^(r*c) twist layer, halving steps of the truncated
 * transforms, the first layers when w*n1 < 128) take the general path: rotated limb reads with a
 * funnel shift, bit-region complement, and the source carry words re-injected at their rotated
 * bit position as one more addend of the chunk's chain.
 */
#ifndef MFFT_TILES_H
#define MFFT_TILES_H

/* (r0,r1) = (a0,a1) + (b0,b1) + cin, cin in {0,1}; returns the carry out */
__device__ __forceinline__ uint32_t add2c(limb_t &r0, limb_t &r1, limb_t a0, limb_t a1, limb_t b0, limb_t b1, uint32_t cin)
{
#ifdef MFFT_EMU
   const mfft_u128 s0 = (mfft_u128) a0 + b0 + (cin ? 1u : 0u);
   const mfft_u128 s1 = (mfft_u128) a1 + b1 + (limb_t)(s0 >> 64);
   r0 = (limb_t) s0; r1 = (limb_t) s1;
   return (uint32_t)(s1 >> 64);
#else
   uint32_t cout, tmp;
   asm("add.cc.u32 %3, %8, 0xffffffff;\n\t"
       "addc.cc.u64 %0, %4, %6;\n\t"
       "addc.cc.u64 %1, %5, %7;\n\t"
       "addc.u32 %2, 0, 0;"
       : "=l"(r0), "=l"(r1), "=r"(cout), "=r"(tmp)
       : "l"(a0), "l"(a1), "l"(b0), "l"(b1), "r"(cin));
   (void) tmp;
   return cout;
#endif
}

/* (r0,r1) = (a0,a1) + (b0,b1); returns the carry out */
__device__ __forceinline__ int32_t add2(limb_t &r0, limb_t &r1, limb_t a0, limb_t a1, limb_t b0, limb_t b1)
{
#ifdef MFFT_EMU
   const mfft_u128 s0 = (mfft_u128) a0 + b0;
   const mfft_u128 s1 = (mfft_u128) a1 + b1 + (limb_t)(s0 >> 64);
   r0 = (limb_t) s0; r1 = (limb_t) s1;
   return (int32_t)(s1 >> 64);
#else
   int32_t cout;
   asm("add.cc.u64 %0, %3, %5;\n\t"
       "addc.cc.u64 %1, %4, %6;\n\t"
       "addc.u32 %2, 0, 0;"
       : "=l"(r0), "=l"(r1), "=r"(cout) : "l"(a0), "l"(a1), "l"(b0), "l"(b1));
   return cout;
#endif
}

/* (r0,r1) = (a0,a1) - (b0,b1) mod B^2; returns -borrow (0 or -1) */
__device__ __forceinline__ int32_t sub2(limb_t &r0, limb_t &r1, limb_t a0, limb_t a1, limb_t b0, limb_t b1)
{
#ifdef MFFT_EMU
   const limb_t d0 = a0 - b0; const uint32_t bw0 = a0 < b0;
   const limb_t d1 = a1 - b1 - bw0; const uint32_t bw1 = (a1 < b1) || (a1 == b1 && bw0);
   r0 = d0; r1 = d1;
   return -(int32_t) bw1;
#else
   int32_t nb;
   asm("sub.cc.u64 %0, %3, %5;\n\t"
       "subc.cc.u64 %1, %4, %6;\n\t"
       "subc.u32 %2, 0, 0;"
       : "=l"(r0), "=l"(r1), "=r"(nb) : "l"(a0), "l"(a1), "l"(b0), "l"(b1));
   return nb;
#endif
}

/* 16-byte chunk accesses (chunk bodies are 16-byte aligned in shared memory) */
__device__ __forceinline__ void ld2(limb_t &x0, limb_t &x1, const limb_t *p)
{
#ifdef MFFT_EMU
   x0 = p[0]; x1 = p[1];
#else
   const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p); x0 = v.x; x1 = v.y;
#endif
}
__device__ __forceinline__ void st2(limb_t *p, limb_t x0, limb_t x1)
{
#ifdef MFFT_EMU
   p[0] = x0; p[1] = x1;
#else
   *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(x0, x1);
#endif
}

template <int NT> struct tile_cfg {
   static constexpr uint32_t L = 64u * NT;           /* body limbs */
   static constexpr uint32_t NCH = 32u * NT;         /* chunks */
   static constexpr uint32_t CW = L + 2;             /* limb offset of the carry words: behind the block image (body, top limb, pad) */
   static constexpr uint32_t SP = CW + NCH / 2;      /* limbs per coefficient: block image (one 16-byte aligned pitch of the slab,
                                                        so that a bulk copy of a whole block lands in place) + int32 carry words */
};

/* one term of an op after folding e >= NW into the sign */
struct tterm {
   uint32_t present, neg;   /* neg: the term is subtracted */
   uint32_t y, bs;          /* rotation = 64*y + bs bits */
   uint32_t yc, yr;         /* y = 2*yc + yr */
};

__device__ __forceinline__ void tterm_setup(tterm &t, int sign, uint32_t e, uint32_t NW)
{
   t.present = (sign != 0); t.neg = 0; t.y = 0; t.bs = 0; t.yc = 0; t.yr = 0;
   if (!sign) return;
   if (e >= NW) { e -= NW; sign = -sign; }
   t.neg = (sign < 0);
   t.y = e >> 6; t.bs = e & 63;
   t.yc = t.y >> 1; t.yr = t.y & 1;
}

/* ---- aligned path ------------------------------------------------------------------------- */
/* Output chunk (ia + ta.yc) of  +-A*B^(2*ycA) +- Bt*B^(2*ycB)  from A's chunk `ia` (a0,a1,ca) and
 * B's chunk `jb` (b0,b1,cb); stores the chunk and its carry word into `out`. */
template <int NT>
__device__ __forceinline__ void aligned_out(limb_t *out, const tterm &ta, const tterm &tb, uint32_t ia, uint32_t jb,
                                            limb_t a0, limb_t a1, int32_t ca, limb_t b0, limb_t b1, int32_t cb)
{
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH;
   uint32_t och = ia + ta.yc, na = ta.neg;
   if (och >= NCH) { och -= NCH; na ^= 1u; }
   const limb_t ma = (limb_t) 0 - (limb_t) na;
   limb_t r0, r1; int32_t k;
   if (tb.present)
   {
      const uint32_t nb = tb.neg ^ ((jb + tb.yc >= NCH) ? 1u : 0u);
      const limb_t mb = (limb_t) 0 - (limb_t) nb;
      uint32_t c = add2c(r0, r1, a0 ^ ma, a1 ^ ma, b0 ^ mb, b1 ^ mb, na | nb);
      k = (int32_t) c - (int32_t)(na + nb) + (na ? -ca : ca) + (nb ? -cb : cb);
      if (na & nb)
      {  /* -(a+b): a second +1 (both operands were complemented) */
         c = add2c(r0, r1, r0, r1, 0, 0, 1u);
         k += (int32_t) c;
      }
   } else
   {
      const uint32_t c = add2c(r0, r1, a0 ^ ma, a1 ^ ma, 0, 0, na);
      k = (int32_t) c - (int32_t) na + (na ? -ca : ca);
   }
#ifdef MFFT_EMU
   out[2 * och] = r0; out[2 * och + 1] = r1;
#else
   *reinterpret_cast<ulonglong2 *>(out + 2 * och) = make_ulonglong2(r0, r1);
#endif
   reinterpret_cast<int32_t *>(out + tile_cfg<NT>::CW)[och] = k;
}

/* ---- general path ------------------------------------------------------------------------- */
/* limbs (2ch, 2ch+1) of the rotated body of X with the negated bit region complemented (bits
 * [0,e) for a positive term, [e,NW) for a negative one), and the addend (e0,e1) + kadj*B^2 = the
 * small signed constant that lands at local limb t.yr, bit t.bs of this chunk: the rotated carry
 * word of source chunk ch-1-yc (negated if it wrapped), and for ch == yc the -+1 at bit e that
 * completes the complement (-F = ~F + 2^a - 2^b; the matching +1 at bit 0 is added by the caller). */
template <int NT>
__device__ __forceinline__ void general_term(limb_t &x0, limb_t &x1, limb_t &e0, limb_t &e1, int32_t &kadj,
                                             const tterm &t, const limb_t *X, uint32_t ch)
{
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH;
   const uint32_t k0 = 2 * ch;
   uint32_t q0 = k0 + L - t.y;
   if (q0 >= L) q0 -= L;
   const uint32_t q1 = (q0 + 1 == L) ? 0 : q0 + 1;
   const limb_t l0 = X[q0], l1 = X[q1];
   if (t.bs)
   {
      const limb_t lm = X[q0 ? q0 - 1 : L - 1];
      const uint32_t rs = 64 - t.bs;
      x0 = (l0 << t.bs) | (lm >> rs);
      x1 = (l1 << t.bs) | (l0 >> rs);
   } else { x0 = l0; x1 = l1; }
   const limb_t nm = (limb_t) 0 - (limb_t) t.neg;
   const limb_t low = (((limb_t) 1 << t.bs) - 1);
   const limb_t m0 = (k0 < t.y) ? ~(limb_t) 0 : ((k0 == t.y) ? low : 0);
   const limb_t m1 = (k0 + 1 < t.y) ? ~(limb_t) 0 : ((k0 + 1 == t.y) ? low : 0);
   x0 ^= m0 ^ nm; x1 ^= m1 ^ nm;
   const uint32_t wrapped = (ch <= t.yc);
   const uint32_t j = wrapped ? ch + NCH - 1 - t.yc : ch - 1 - t.yc;
   int64_t c = (int64_t) reinterpret_cast<const int32_t *>(X + tile_cfg<NT>::CW)[j];
   if (wrapped) c = -c;
   if (ch == t.yc) c -= 1;
   const int64_t V = t.neg ? -c : c;
   const limb_t lo = (limb_t) V << t.bs;
   const int64_t hi = t.bs ? (V >> (64 - t.bs)) : (V >> 63);
   if (t.yr == 0) { e0 = lo; e1 = (limb_t) hi; kadj = (int32_t)(hi >> 63); }
   else { e0 = 0; e1 = lo; kadj = (int32_t) hi; }
}

/* chunk ch of  +-A*2^eA +- B*2^eB  into (r0,r1), returns the chunk's carry word */
template <int NT>
__device__ __forceinline__ int32_t general_out(limb_t &r0, limb_t &r1, const tterm &ta, const limb_t *A, const tterm &tb,
                                               const limb_t *B, uint32_t ch)
{
   /* the +1 at bit 0 of each term enters as carry-ins of the chains of chunk 0 */
   const uint32_t ones = (ch == 0) ? ta.present + tb.present : 0u;
   limb_t xa0 = 0, xa1 = 0, ea0 = 0, ea1 = 0, xb0 = 0, xb1 = 0, eb0 = 0, eb1 = 0;
   int32_t ka = 0, kb = 0, k;
   if (ta.present) general_term<NT>(xa0, xa1, ea0, ea1, ka, ta, A, ch);
   if (tb.present) general_term<NT>(xb0, xb1, eb0, eb1, kb, tb, B, ch);
   k = ka + kb;
   k += (int32_t) add2c(r0, r1, xa0, xa1, xb0, xb1, ones >= 1u);
   if (ta.present) k += (int32_t) add2c(r0, r1, r0, r1, ea0, ea1, ones >= 2u);
   if (tb.present) k += (int32_t) add2c(r0, r1, r0, r1, eb0, eb1, 0u);
   return k;
}

/* ---- leaving the tile: resolve the carry words --------------------------------------------- */
/* The warp turns the carry-save coefficient at sblk into plain limbs (written back to the body)
 * and returns the signed top limb.  Lane j owns chunks [j*NT, j*NT+NT). */
template <int NT>
__device__ __forceinline__ int64_t resolve_carries(limb_t *sblk, uint32_t lane)
{
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH;
   const int32_t *cw = reinterpret_cast<const int32_t *>(sblk + tile_cfg<NT>::CW);
   limb_t r[2 * NT];
   int64_t cy = 0;
#pragma unroll
   for (int t = 0; t < NT; t++)
   {
      const uint32_t ch = lane * NT + t;
      if (ch) cy += (int64_t) cw[ch - 1];
#pragma unroll
      for (int i = 0; i < 2; i++)
      {
         const mfft_i128 acc = (mfft_i128) cy + (mfft_i128)(mfft_u128) sblk[2 * ch + i];
         r[2 * t + i] = (limb_t) acc; cy = (int64_t)(acc >> 64);
      }
   }
   int64_t d = cy;                                     /* carry into the next lane's first chunk */
   int64_t top = (int64_t) cw[NCH - 1];
   for (;;)
   {
      int64_t din = __shfl_up_sync(FULL, d, 1);
      if (lane == 0) din = 0;
      top += __shfl_sync(FULL, d, 31);
      d = 0;
      if (!__any_sync(FULL, din != 0)) break;
      if (din != 0)
      {
         int64_t c = din;
#pragma unroll
         for (int i = 0; i < 2 * NT; i++)
         {
            const mfft_i128 acc = (mfft_i128) c + (mfft_i128)(mfft_u128) r[i];
            r[i] = (limb_t) acc; c = (int64_t)(acc >> 64);
         }
         d = c;
      }
   }
   __syncwarp();
#pragma unroll
   for (int i = 0; i < 2 * NT; i++) sblk[lane * 2 * NT + i] = r[i];
   __syncwarp();
   return top;
}

/* d[ti] in {-1,0,1}: what still has to be added to the chunk after ti*32+lane.  Moves the carries
 * on and returns what leaves the last chunk.  Almost always nothing moves (a carry travels on only
 * through an all-zero / all-one chunk); otherwise the chain is solved exactly in a fixed number of
 * steps as a prefix scan of three-state transition functions: with `in` the carry entering a chunk,
 *      out(in) = d + [in = +1 and chunk all ones] - [in = -1 and chunk zero]      (in {-1,0,1})
 * (sparse or cancelling coefficients -- e.g. the exact zeros beyond the end of a product -- would
 * otherwise ripple across the whole coefficient one chunk per round). */
__device__ __forceinline__ uint32_t tf_apply(uint32_t F, uint32_t st) { return (F >> (2u * st)) & 3u; }
__device__ __forceinline__ uint32_t tf_compose(uint32_t G, uint32_t F)      /* G after F */
{ return tf_apply(G, tf_apply(F, 0)) | (tf_apply(G, tf_apply(F, 1)) << 2) | (tf_apply(G, tf_apply(F, 2)) << 4); }

template <int NT>
__device__ __forceinline__ int64_t ripple_regs(limb_t (&r0)[NT], limb_t (&r1)[NT], int32_t (&d)[NT], uint32_t lane)
{
   {  /* fast check: nothing to hand on except out of the very last chunk */
      bool any = false;
#pragma unroll
      for (int ti = 0; ti < NT; ti++) any = any || (d[ti] != 0 && !(ti == NT - 1 && lane == 31));
      if (!__any_sync(FULL, any)) return (int64_t) __shfl_sync(FULL, d[NT - 1], 31);
   }
   uint32_t state = 1;                                  /* carry entering the row, biased by 1 */
#pragma unroll
   for (int ti = 0; ti < NT; ti++)
   {
      const uint32_t ones = ((r0[ti] & r1[ti]) == ~(limb_t) 0), zero = ((r0[ti] | r1[ti]) == 0);
      const int32_t dd = d[ti];
      /* outputs for in = -1, 0, +1, biased by 1 */
      uint32_t F = (uint32_t)(dd - (int32_t) zero + 1) | ((uint32_t)(dd + 1) << 2) | ((uint32_t)(dd + (int32_t) ones + 1) << 4);
#pragma unroll
      for (int off = 1; off < 32; off <<= 1)
      {
         const uint32_t Fp = __shfl_up_sync(FULL, F, off);
         if (lane >= (uint32_t) off) F = tf_compose(F, Fp);
      }
      uint32_t Fex = __shfl_up_sync(FULL, F, 1);        /* everything before this lane's chunk */
      const uint32_t in = lane ? tf_apply(Fex, state) : state;
      state = tf_apply(__shfl_sync(FULL, F, 31), state);
      const int64_t c = (int64_t) in - 1;
      (void) add2(r0[ti], r1[ti], r0[ti], r1[ti], (limb_t) c, (limb_t)(c >> 63));
   }
   return (int64_t) state - 1;
}

/* The same without touching shared memory again: lane j ends up with the resolved limbs of chunks
 * ti*32 + j in (r0[ti], r1[ti]) -- 16 consecutive bytes per lane, i.e. one coalesced 512-byte
 * store per ti.  Absorbing the previous chunk's carry word leaves a carry in {-1,0,1}, which
 * moves on only through all-zero / all-one chunks, so the ripple loop almost never iterates. */
template <int NT>
__device__ __forceinline__ int64_t resolve_regs(limb_t (&r0)[NT], limb_t (&r1)[NT], const limb_t *sblk, uint32_t lane)
{
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH;
   const int32_t *cw = reinterpret_cast<const int32_t *>(sblk + tile_cfg<NT>::CW);
   int32_t d[NT];
#pragma unroll
   for (int ti = 0; ti < NT; ti++)
   {
      const uint32_t i = ti * 32u + lane;
      limb_t x0, x1;
      ld2(x0, x1, sblk + 2 * i);
      const int32_t c = i ? cw[i - 1] : 0;
      const int32_t k = add2(r0[ti], r1[ti], x0, x1, (limb_t)(int64_t) c, (limb_t)((int64_t) c >> 63));
      d[ti] = k + (c >> 31);
   }
   return (int64_t) cw[NCH - 1] + ripple_regs<NT>(r0, r1, d, lane);
}

/* mpn_normmod_2expp1 (mul_fft.c:272-294) on resolved limbs held in registers (chunk ti*32+lane in
 * (r0[ti], r1[ti])): value = body + top*B^l == body - top */
template <int NT>
__device__ __forceinline__ int64_t normalise_regs(limb_t (&r0)[NT], limb_t (&r1)[NT], int64_t top, uint32_t lane)
{
   for (int it = 0; it < 4; it++)
   {
      if (top == 0) break;
      if (top == 1)
      {
         bool z = true;
#pragma unroll
         for (int ti = 0; ti < NT; ti++) z = z && ((r0[ti] | r1[ti]) == 0);
         if (__all_sync(FULL, z)) break;
      }
      int32_t d[NT];
#pragma unroll
      for (int ti = 0; ti < NT; ti++) d[ti] = 0;
      if (lane == 0)
      {  /* body -= top at limb 0 */
         const int64_t c = -top;
         const int32_t k = add2(r0[0], r1[0], r0[0], r1[0], (limb_t) c, (limb_t)(c >> 63));
         d[0] = k + (int32_t)(c >> 63);
      }
      top = ripple_regs<NT>(r0, r1, d, lane);
   }
   return top;
}

/* block of physical position pos (< S, slab half 0) of batch entry b */
__device__ __forceinline__ limb_t *tile_block_ptr(limb_t *slab, const mfft_geom &g, uint32_t pos, const mfft_batch &b)
{
   const uint64_t idx = (uint64_t) b.parity * g.half_blocks + b.base + (uint64_t) pos * g.slot_stride;
   return slab + idx * g.pitch;
}

/* ---- bulk asynchronous copies (TMA engine, cp.async.bulk -> UBLKCP) completing on an mbarrier ---- */
/* A tile is fetched by a handful of bulk copies -- one whole (l+2)-limb block image per coefficient
 * plus the tile's op descriptors -- issued by the lanes of one warp; the copy engine moves the data
 * while no thread of the CTA spends issue slots on it, and every thread then waits on the barrier's
 * phase.  MBAR_BYTES of shared memory are reserved for the barrier object. */
#define MBAR_BYTES 32
#ifdef MFFT_EMU
typedef emu_mbar tile_mbar;
__device__ __forceinline__ void mbar_init(tile_mbar *m, uint32_t count) { emu_mbar_init(m, count); }
__device__ __forceinline__ void mbar_arrive_expect_tx(tile_mbar *m, uint32_t bytes) { emu_mbar_arrive_expect_tx(m, bytes); }
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, tile_mbar *m) { emu_bulk_copy(sdst, gsrc, bytes, m); }
__device__ __forceinline__ void mbar_wait(tile_mbar *m, uint32_t parity) { emu_mbar_wait(m, parity); }
__device__ __forceinline__ void fence_async_smem() { }
#else
typedef unsigned long long tile_mbar;
__device__ __forceinline__ void mbar_init(tile_mbar *m, uint32_t count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t) __cvta_generic_to_shared(m)), "r"(count) : "memory");
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(tile_mbar *m, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                :: "r"((uint32_t) __cvta_generic_to_shared(m)), "r"(bytes) : "memory");
}
/* bytes: a multiple of 16; both addresses 16-byte aligned */
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, tile_mbar *m)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"((uint32_t) __cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes),
                   "r"((uint32_t) __cvta_generic_to_shared(m)) : "memory");
}
__device__ __forceinline__ void mbar_wait(tile_mbar *m, uint32_t parity)
{
   asm volatile("{\n\t.reg .pred p;\n\t"
                "MBAR_WAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra MBAR_DONE_%=;\n\t"
                "bra MBAR_WAIT_%=;\n\t"
                "MBAR_DONE_%=:\n\t}"
                :: "r"((uint32_t) __cvta_generic_to_shared(m)), "r"(parity) : "memory");
}
/* orders earlier generic-proxy accesses to shared memory before later asynchronous-proxy ones */
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

/* 16-byte asynchronous copy global -> shared (LDGSTS), L2 only: the tile is read exactly once */
__device__ __forceinline__ void cp_async16(limb_t *sdst, const limb_t *gsrc)
{
#ifdef MFFT_EMU
   sdst[0] = gsrc[0]; sdst[1] = gsrc[1];
#else
   const uint32_t sa = (uint32_t) __cvta_generic_to_shared(sdst);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8(limb_t *sdst, const limb_t *gsrc)
{
#ifdef MFFT_EMU
   sdst[0] = gsrc[0];
#else
   const uint32_t sa = (uint32_t) __cvta_generic_to_shared(sdst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4(int32_t *sdst, const int32_t *gsrc)
{
#ifdef MFFT_EMU
   sdst[0] = gsrc[0];
#else
   const uint32_t sa = (uint32_t) __cvta_generic_to_shared(sdst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all()
{
#ifndef MFFT_EMU
   asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
}

/* warp-level mpn_normmod_2expp1 (mul_fft.c:272-294) on resolved limbs */
__device__ __forceinline__ int64_t normalise_tile(limb_t *sblk, uint32_t l, int64_t top, uint32_t lane)
{
   for (int it = 0; it < 4; it++)
   {
      if (top == 0) break;
      if (top == 1)
      {
         bool z = true;
         for (uint32_t k = lane; k < l; k += 32) z = z && (sblk[k] == 0);
         if (__all_sync(FULL, z)) break;
      }
      int64_t nt = 0;
      if (lane == 0)
      {  /* body -= top, what leaves limb l-1 is the new top */
         int64_t c = -top;
         for (uint32_t pos = 0; c != 0; pos++)
         {
            if (pos == l) { nt = c; break; }
            const mfft_i128 acc = (mfft_i128) c + (mfft_i128)(mfft_u128) sblk[pos];
            sblk[pos] = (limb_t) acc; c = (int64_t)(acc >> 64);
         }
      }
      __syncwarp();
      top = __shfl_sync(FULL, nt, 0);
   }
   return top;
}

/* rendezvous of the W warps that share an op (between "everybody has read its operand chunks" and the
   in-place stores): the warp itself, or a named barrier of the W consecutive warps */
template <int W>
__device__ __forceinline__ void op_sync(uint32_t warp)
{
   if (W == 1) { __syncwarp(); return; }
#ifdef MFFT_EMU
   emu_bar_sync(1 + warp / W, 32 * W);
#else
   asm volatile("bar.sync %0, %1;" :: "r"(1u + warp / W), "r"(32u * W) : "memory");
#endif
}

/* ---- radix-4 units: two radix-2 layers per shared-memory round trip -------------------------- */
/* signed chunk value = 2 limbs + carry word */
struct cval { limb_t x0, x1; int32_t c; };
__device__ __forceinline__ cval cv_add(const cval &a, const cval &b)
{ cval r; r.c = add2(r.x0, r.x1, a.x0, a.x1, b.x0, b.x1) + a.c + b.c; return r; }
__device__ __forceinline__ cval cv_sub(const cval &a, const cval &b)
{ cval r; r.c = sub2(r.x0, r.x1, a.x0, a.x1, b.x0, b.x1) + a.c - b.c; return r; }
__device__ __forceinline__ cval cv_load(const limb_t *P, uint32_t cwoff, uint32_t ch)      /* cwoff = tile_cfg<NT>::CW */
{ cval r; ld2(r.x0, r.x1, P + 2 * ch); r.c = reinterpret_cast<const int32_t *>(P + cwoff)[ch]; return r; }
__device__ __forceinline__ void cv_store(limb_t *P, uint32_t cwoff, uint32_t ch, const cval &v)
{ st2(P + 2 * ch, v.x0, v.x1); reinterpret_cast<int32_t *>(P + cwoff)[ch] = v.c; }

/* Two forward layers on positions P0..P3 in place (FFT_radix2 786-827, two levels of the recursion):
 *    layer 1:  (P0,P2) -> P0+P2, +-(P0-P2) 2^(128 y1)      (P1,P3) -> P1+P3, +-(P1-P3) 2^(128 y1')
 *    layer 2:  (P0,P1) -> ...,   2^(128 y2)                  (P2,P3) -> ...,   2^(128 y2')
 * with y1' = y1 + NCH/2 (the two layer-1 twiddles differ by the quarter turn 2^(NW/2)).  A rotation by
 * NCH/2 = 16 NT chunks maps a lane's chunk ti to its own chunk ti +- NT/2, so the whole unit is
 * lane-local: 4 chunk loads and 4 chunk stores per lane and ti instead of 8 + 8.  q[k] = y | neg << 31. */
template <int NT, int W>
__device__ __forceinline__ void fwd4_unit(limb_t *P0, limb_t *P1, limb_t *P2, limb_t *P3,
                                          uint32_t q1, uint32_t q1p, uint32_t q2, uint32_t q2p, uint32_t lane, uint32_t warp)
{
   constexpr uint32_t NCH = tile_cfg<NT>::NCH; constexpr int NTS = NT / W; const uint32_t half = warp % W;
   const uint32_t y1 = q1 & 0x7fffffffu, y1p = q1p & 0x7fffffffu, y2 = q2 & 0x7fffffffu, y2p = q2p & 0x7fffffffu;
   const uint32_t n1 = q1 >> 31, n1p = q1p >> 31, n2 = q2 >> 31, n2p = q2p >> 31;
   cval a0[NTS], a1[NTS], a2[NTS], a3[NTS];
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t i = (half + W * tj) * 32u + lane;
      a0[tj] = cv_load(P0, tile_cfg<NT>::CW, i); a1[tj] = cv_load(P1, tile_cfg<NT>::CW, i); a2[tj] = cv_load(P2, tile_cfg<NT>::CW, i); a3[tj] = cv_load(P3, tile_cfg<NT>::CW, i);
   }
   op_sync<W>(warp);
   /* layer 1 in place: a0 <- a0+a2, a2 <- +-(a0-a2) (the chunk that lands at i+y1), same for a1,a3 */
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t i = (half + W * tj) * 32u + lane;
      const cval s = cv_add(a0[tj], a2[tj]);
      a2[tj] = (((i + y1 >= NCH) ? 1u : 0u) != n1) ? cv_sub(a2[tj], a0[tj]) : cv_sub(a0[tj], a2[tj]);
      a0[tj] = s;
      const cval sp = cv_add(a1[tj], a3[tj]);
      a3[tj] = (((i + y1p >= NCH) ? 1u : 0u) != n1p) ? cv_sub(a3[tj], a1[tj]) : cv_sub(a1[tj], a3[tj]);
      a1[tj] = sp;
   }
   /* layer 2 */
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t i = (half + W * tj) * 32u + lane;
      constexpr int dummy = 0; (void) dummy;
      const int tp = (tj + NTS / 2) % NTS;                        /* a3's chunk that lands where a2[tj] does */
      cv_store(P0, tile_cfg<NT>::CW, i, cv_add(a0[tj], a1[tj]));
      {
         uint32_t o = i + y2, n = n2;
         if (o >= NCH) { o -= NCH; n ^= 1u; }
         cv_store(P1, tile_cfg<NT>::CW, o, n ? cv_sub(a1[tj], a0[tj]) : cv_sub(a0[tj], a1[tj]));
      }
      uint32_t j = i + y1; if (j >= NCH) j -= NCH;
      cv_store(P2, tile_cfg<NT>::CW, j, cv_add(a2[tj], a3[tp]));
      {
         uint32_t o = j + y2p, n = n2p;
         if (o >= NCH) { o -= NCH; n ^= 1u; }
         cv_store(P3, tile_cfg<NT>::CW, o, n ? cv_sub(a3[tp], a2[tj]) : cv_sub(a2[tj], a3[tp]));
      }
   }
}

/* Two inverse layers on positions P0..P3 in place (IFFT_radix2 1444-1536, two levels):
 *    layer 1:  (P0,P1) -> P0 +- P1 2^(128 y2), P0 -+ ...      (P2,P3) -> P2 +- P3 2^(128 y2'), ...
 *    layer 2:  (P0,P2) -> P0 +- P2 2^(128 y1), ...            (P1,P3) -> P1 +- P3 2^(128 y1'), ...
 * with y1' = y1 + NCH/2; exponents are the negated (inverse) rotations already folded mod 2NW. */
template <int NT, int W>
__device__ __forceinline__ void inv4_unit(limb_t *P0, limb_t *P1, limb_t *P2, limb_t *P3,
                                          uint32_t q2, uint32_t q2p, uint32_t q1, uint32_t q1p, uint32_t lane, uint32_t warp)
{
   constexpr uint32_t NCH = tile_cfg<NT>::NCH; constexpr int NTS = NT / W; const uint32_t half = warp % W;
   const uint32_t y1 = q1 & 0x7fffffffu, y1p = q1p & 0x7fffffffu, y2 = q2 & 0x7fffffffu, y2p = q2p & 0x7fffffffu;
   const uint32_t n1 = q1 >> 31, n1p = q1p >> 31, n2 = q2 >> 31, n2p = q2p >> 31;
   cval u0[NTS], u1[NTS], u2[NTS], u3[NTS];
   /* u0,u1 at chunk i; u2,u3 at chunk m = i - y1 (what layer 2 adds to chunk i of P0) */
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t i = (half + W * tj) * 32u + lane;
      const uint32_t jb = (i >= y2) ? i - y2 : i + NCH - y2;
      const uint32_t m = (i >= y1) ? i - y1 : i + NCH - y1;
      const uint32_t mb = (m >= y2p) ? m - y2p : m + NCH - y2p;
      u0[tj] = cv_load(P0, tile_cfg<NT>::CW, i); u1[tj] = cv_load(P1, tile_cfg<NT>::CW, jb); u2[tj] = cv_load(P2, tile_cfg<NT>::CW, m); u3[tj] = cv_load(P3, tile_cfg<NT>::CW, mb);
   }
   op_sync<W>(warp);
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t i = (half + W * tj) * 32u + lane;
      const uint32_t m = (i >= y1) ? i - y1 : i + NCH - y1;
      {  /* (u0,u1) <- b0 +- b1', b0 -+ b1' */
         const cval sm = cv_add(u0[tj], u1[tj]), df = cv_sub(u0[tj], u1[tj]);
         const bool ng = ((i < y2) ? 1u : 0u) != n2;
         u0[tj] = ng ? df : sm; u1[tj] = ng ? sm : df;
      }
      {
         const cval sm = cv_add(u2[tj], u3[tj]), df = cv_sub(u2[tj], u3[tj]);
         const bool ng = ((m < y2p) ? 1u : 0u) != n2p;
         u2[tj] = ng ? df : sm; u3[tj] = ng ? sm : df;
      }
   }
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t i = (half + W * tj) * 32u + lane;
      const int tp = (tj + NTS / 2) % NTS;                        /* u3 at chunk i - y1' = (i -+ NCH/2) - y1 */
      {
         const bool ng = ((i < y1) ? 1u : 0u) != n1;       /* P2's chunk enters negated */
         limb_t *ds = ng ? P2 : P0, *dt = ng ? P0 : P2;
         cv_store(ds, tile_cfg<NT>::CW, i, cv_add(u0[tj], u2[tj]));
         cv_store(dt, tile_cfg<NT>::CW, i, cv_sub(u0[tj], u2[tj]));
      }
      {
         const bool ng = ((i < y1p) ? 1u : 0u) != n1p;
         limb_t *ds = ng ? P3 : P1, *dt = ng ? P1 : P3;
         cv_store(ds, tile_cfg<NT>::CW, i, cv_add(u1[tj], u3[tp]));
         cv_store(dt, tile_cfg<NT>::CW, i, cv_sub(u1[tj], u3[tp]));
      }
   }
}

/* ---- one-operand rotation by any number of bits, chunk-local --------------------------------- */
/* S = +-A * 2^t, 0 <= t < NW.  Written as a rotation by yc1 = ceil(t/128) whole chunks (chunks that
 * wrap around p are negated) followed by a right shift by s = 128 yc1 - t < 128 bits.  The shift
 * is chunk-local in carry-save form: the s bits leaving chunk j+1 -- and the low s bits of chunk
 * j's own carry word -- are worth 2^(128-s) in chunk j; what leaves chunk 0 re-enters negated at
 * the top (2^-s == -2^(NW-s)).  Used for the z^(r*c) twist rotations (1409-1411), whose exponents
 * depend on the column. */
__device__ __forceinline__ void shr128(limb_t &h0, limb_t &h1, limb_t x0, limb_t x1, uint32_t s)   /* 1 <= s <= 127 */
{
   if (s >= 64) { h0 = x1 >> (s - 64); h1 = 0; }
   else { h0 = (x0 >> s) | (x1 << (64 - s)); h1 = x1 >> s; }
}
__device__ __forceinline__ void shl128(limb_t &h0, limb_t &h1, limb_t x0, limb_t x1, uint32_t u)   /* 1 <= u <= 127 */
{
   if (u >= 64) { h1 = x0 << (u - 64); h0 = 0; }
   else { h1 = (x1 << u) | (x0 >> (64 - u)); h0 = x0 << u; }
}

template <int NT, int W>
__device__ __forceinline__ void rotg_unit(limb_t *S, const limb_t *A, uint32_t t, uint32_t neg, uint32_t lane, uint32_t warp)
{
   constexpr uint32_t NCH = tile_cfg<NT>::NCH; constexpr int NTS = NT / W; const uint32_t half = warp % W;
   uint32_t yc1 = (t + 127u) >> 7;
   const uint32_t s = 128u * yc1 - t;                       /* 0..127 */
   if (yc1 == NCH) { yc1 = 0; neg ^= 1u; }                  /* a full turn is a factor -1 */
   const int32_t *cwA = reinterpret_cast<const int32_t *>(A + tile_cfg<NT>::CW);
   cval r[NTS]; limb_t y0[NTS], y1[NTS];
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t j = (half + W * tj) * 32u + lane;
      {  /* R_j = +-A_(j - yc1) */
         const uint32_t wj = (j < yc1) ? 1u : 0u, sj = wj ? j + NCH - yc1 : j - yc1;
         ld2(r[tj].x0, r[tj].x1, A + 2 * sj); r[tj].c = cwA[sj];
         if (wj != neg) { const int32_t b = sub2(r[tj].x0, r[tj].x1, 0, 0, r[tj].x0, r[tj].x1); r[tj].c = b - r[tj].c; }
      }
      {  /* body of R_(j+1); for the last chunk that is R_0, which enters negated (handled below) */
         const uint32_t j1 = (j + 1 == NCH) ? 0u : j + 1;
         const uint32_t w1 = (j1 < yc1) ? 1u : 0u, s1 = w1 ? j1 + NCH - yc1 : j1 - yc1;
         ld2(y0[tj], y1[tj], A + 2 * s1);
         if (w1 != neg) (void) sub2(y0[tj], y1[tj], 0, 0, y0[tj], y1[tj]);
      }
   }
   op_sync<W>(warp);
   int32_t *cwS = reinterpret_cast<int32_t *>(S + tile_cfg<NT>::CW);
#pragma unroll
   for (int tj = 0; tj < NTS; tj++)
   {
      const uint32_t j = (half + W * tj) * 32u + lane;
      if (s == 0) { st2(S + 2 * j, r[tj].x0, r[tj].x1); cwS[j] = r[tj].c; continue; }
      limb_t h0, h1, f0, f1, g0, g1, o0, o1;
      const int64_t c64 = (int64_t) r[tj].c;
      shr128(h0, h1, r[tj].x0, r[tj].x1, s);
      shl128(f0, f1, (limb_t) c64, (limb_t)(c64 >> 63), 128u - s);
      shl128(g0, g1, y0[tj], y1[tj], 128u - s);
      h0 |= f0; h1 |= f1;                                   /* disjoint bit fields */
      const int32_t k = (j + 1 == NCH) ? sub2(o0, o1, h0, h1, g0, g1) : add2(o0, o1, h0, h1, g0, g1);
      st2(S + 2 * j, o0, o1);
      cwS[j] = ((s >= 31) ? (r[tj].c >> 31) : (r[tj].c >> s)) + k;
   }
}

/* barrier of the threads that work on one tile: the whole CTA (bar_id 0) or one warp group of a
   persistent CTA (named barrier bar_id, bar_threads threads) */
__device__ __forceinline__ void tile_sync(uint32_t bar_id, uint32_t bar_threads)
{
#ifdef MFFT_EMU
   if (bar_id == 0) __syncthreads(); else emu_bar_sync(bar_id, bar_threads);
#else
   if (bar_id == 0) __syncthreads();
   else asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_threads) : "memory");
#endif
}

/* all stages of a tile on the coefficients in shared memory (sops sorted by stage, sst[k] = first op
   of stage k); warp `warp` of `nwarps` takes every nwarps-th op of a stage */
template <int NT, int W>
__device__ __forceinline__ void tile_stages(limb_t *coef, const mfft_tileop *sops, const uint32_t *sst, uint32_t nstages,
                                            const mfft_batch &b, uint32_t warp, uint32_t nwarps, uint32_t lane,
                                            uint32_t bar_id, uint32_t bar_threads)
{
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH, SP = tile_cfg<NT>::SP, CW = tile_cfg<NT>::CW;
   constexpr uint32_t NW = 64u * L, M2 = 2u * NW;
   constexpr uint32_t AMASK = 127u;                           /* exponent bits that break chunk alignment */
   /* W warps share an op: warp `warp` works on the chunks ti*32 + lane with ti = half, half + W, ... */
   constexpr int NTS = NT / W; const uint32_t half = warp % W, wunit = warp / W, nunits = nwarps / W;
   for (uint32_t st = 0; st < nstages; st++)
   {
      const uint32_t o0 = sst[st], o1 = sst[st + 1];
      for (uint32_t oi = o0 + wunit; oi < o1; oi += nunits)
      {
         const mfft_tileop op = sops[oi];
         const limb_t *A = coef + (size_t) op.a * SP;
         const bool hasB = (op.b != 0xFFFF), hasT = (op.t != 0xFFFF);
         const limb_t *B = hasB ? coef + (size_t) op.b * SP : A;
         limb_t *S = coef + (size_t) op.s * SP;
         limb_t *Tt = hasT ? coef + (size_t) op.t * SP : S;
         if (op.kind == MFFT_K_FWD4 || op.kind == MFFT_K_INV4)
         {  /* two layers at once on four positions (a, b, s, t hold them; the exponent fields the
               four packed rotations); only built for even NT <= 4 */
            if (NT <= 4 && NT % 2 == 0)
            {
               limb_t *P0 = coef + (size_t) op.a * SP, *P1 = coef + (size_t) op.b * SP;
               limb_t *P2 = coef + (size_t) op.s * SP, *P3 = coef + (size_t) op.t * SP;
               if (op.kind == MFFT_K_FWD4) fwd4_unit<(NT <= 4 && NT % 2 == 0) ? NT : 2, (NT <= 4 && NT % 2 == 0) ? W : 1>(P0, P1, P2, P3, op.eSA, op.eSB, op.eTA, op.eTB, lane, warp);
               else inv4_unit<(NT <= 4 && NT % 2 == 0) ? NT : 2, (NT <= 4 && NT % 2 == 0) ? W : 1>(P0, P1, P2, P3, op.eSA, op.eSB, op.eTA, op.eTB, lane, warp);
            }
            continue;
         }
         if (op.kind != MFFT_K_ANY)
         {  /* host-classified aligned shapes: lane-local chunk arithmetic, nothing to decode */
            const uint32_t yc = op.kparam & 0x7fffffffu, neg = op.kparam >> 31;
            const int32_t *cwA = reinterpret_cast<const int32_t *>(A + tile_cfg<NT>::CW);
            const int32_t *cwB = reinterpret_cast<const int32_t *>(B + tile_cfg<NT>::CW);
            int32_t *cwS = reinterpret_cast<int32_t *>(S + tile_cfg<NT>::CW), *cwT = reinterpret_cast<int32_t *>(Tt + tile_cfg<NT>::CW);
            limb_t a0[NTS], a1[NTS], b0[NTS], b1[NTS]; int32_t ca[NTS], cb[NTS];
            if (op.kind == MFFT_K_FWD)
            {
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  ld2(a0[tj], a1[tj], A + 2 * i); ca[tj] = cwA[i];
                  ld2(b0[tj], b1[tj], B + 2 * i); cb[tj] = cwB[i];
               }
               op_sync<W>(warp);
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  limb_t r0, r1; int32_t k;
                  k = add2(r0, r1, a0[tj], a1[tj], b0[tj], b1[tj]);
                  st2(S + 2 * i, r0, r1); cwS[i] = k + ca[tj] + cb[tj];
                  uint32_t o = i + yc, n = neg;
                  if (o >= NCH) { o -= NCH; n ^= 1u; }
                  if (n) k = sub2(r0, r1, b0[tj], b1[tj], a0[tj], a1[tj]) + cb[tj] - ca[tj];
                  else   k = sub2(r0, r1, a0[tj], a1[tj], b0[tj], b1[tj]) + ca[tj] - cb[tj];
                  st2(Tt + 2 * o, r0, r1); cwT[o] = k;
               }
            } else if (op.kind == MFFT_K_INV)
            {
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  const uint32_t j = (i >= yc) ? i - yc : i + NCH - yc;
                  ld2(a0[tj], a1[tj], A + 2 * i); ca[tj] = cwA[i];
                  ld2(b0[tj], b1[tj], B + 2 * j); cb[tj] = cwB[j];
               }
               op_sync<W>(warp);
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  const bool n = ((i < yc) ? 1u : 0u) != neg;      /* B's chunk enters negated */
                  limb_t *ds = n ? Tt : S, *dt = n ? S : Tt;
                  limb_t r0, r1; int32_t k;
                  k = add2(r0, r1, a0[tj], a1[tj], b0[tj], b1[tj]);
                  st2(ds + 2 * i, r0, r1); reinterpret_cast<int32_t *>(ds + tile_cfg<NT>::CW)[i] = k + ca[tj] + cb[tj];
                  k = sub2(r0, r1, a0[tj], a1[tj], b0[tj], b1[tj]);
                  st2(dt + 2 * i, r0, r1); reinterpret_cast<int32_t *>(dt + tile_cfg<NT>::CW)[i] = k + ca[tj] - cb[tj];
               }
            } else if (op.kind == MFFT_K_ROT)
            {
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  ld2(a0[tj], a1[tj], A + 2 * i); ca[tj] = cwA[i];
               }
               op_sync<W>(warp);
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  uint32_t o = i + yc, n = neg;
                  if (o >= NCH) { o -= NCH; n ^= 1u; }
                  limb_t r0 = a0[tj], r1 = a1[tj]; int32_t k = ca[tj];
                  if (n) k = sub2(r0, r1, 0, 0, a0[tj], a1[tj]) - ca[tj];
                  st2(S + 2 * o, r0, r1); cwS[o] = k;
               }
            } else if (op.kind == MFFT_K_2AMB)
            {  /* S = 2A - B [, T = +-(A - B) 2^(128 yc)]  (1630-1631, 1639-1647): chunk-local like the doubling */
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  ld2(a0[tj], a1[tj], A + 2 * i); ca[tj] = cwA[i];
                  ld2(b0[tj], b1[tj], B + 2 * i); cb[tj] = cwB[i];
               }
               op_sync<W>(warp);
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  limb_t r0, r1; int32_t k;
                  const limb_t d0 = a0[tj] << 1, d1 = (a1[tj] << 1) | (a0[tj] >> 63);
                  const int32_t dc = 2 * ca[tj] + (int32_t)(a1[tj] >> 63);
                  k = sub2(r0, r1, d0, d1, b0[tj], b1[tj]);
                  st2(S + 2 * i, r0, r1); cwS[i] = k + dc - cb[tj];
                  if (hasT)
                  {
                     uint32_t o = i + yc, n = neg;
                     if (o >= NCH) { o -= NCH; n ^= 1u; }
                     if (n) k = sub2(r0, r1, b0[tj], b1[tj], a0[tj], a1[tj]) + cb[tj] - ca[tj];
                     else   k = sub2(r0, r1, a0[tj], a1[tj], b0[tj], b1[tj]) + ca[tj] - cb[tj];
                     st2(Tt + 2 * o, r0, r1); cwT[o] = k;
                  }
               }
            } else if (op.kind == MFFT_K_DBL)
            {  /* 2 (x + c B^2) = (2x mod B^2) + (2c + top bit of x) B^2: chunk-local */
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  limb_t x0, x1;
                  ld2(x0, x1, A + 2 * i);
                  const int32_t c = cwA[i];
                  st2(S + 2 * i, x0 << 1, (x1 << 1) | (x0 >> 63));
                  cwS[i] = 2 * c + (int32_t)(x1 >> 63);
               }
            } else if (op.kind == MFFT_K_HALF)
            {  /* (A + B) / 2: halve chunk by chunk; a chunk's (and a carry word's) lowest bit is worth
                  2^127 one chunk below, and chunk 0's lowest bit -2^(NW-1) (2^-1 == -2^(NW-1) mod p) */
               uint32_t lowbit[NTS];
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane, nx = (i + 1 == NCH) ? 0u : i + 1;
                  ld2(a0[tj], a1[tj], A + 2 * i); ca[tj] = cwA[i];
                  ld2(b0[tj], b1[tj], B + 2 * i); cb[tj] = cwB[i];
                  lowbit[tj] = (uint32_t)((A[2 * nx] ^ B[2 * nx]) & 1u);
               }
               op_sync<W>(warp);
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  limb_t t0, t1;
                  const int32_t ct = add2(t0, t1, a0[tj], a1[tj], b0[tj], b1[tj]) + ca[tj] + cb[tj];
                  const int32_t e = (ct & 1) + ((i + 1 == NCH) ? -(int32_t) lowbit[tj] : (int32_t) lowbit[tj]);   /* in {-1,0,1,2} */
                  const limb_t h0 = (t0 >> 1) | (t1 << 63);
                  const limb_t h1 = (t1 >> 1) | ((limb_t)(e & 1) << 63);
                  st2(S + 2 * i, h0, h1);
                  cwS[i] = (ct >> 1) + (e >> 1);
               }
            } else if (op.kind == MFFT_K_SHR)
            {  /* A / 2^s chunk by chunk: the s bits shifted out of a chunk (and of its carry word) are
                  worth 2^(128-s) one chunk below; chunk 0's go to the top negated (2^-s == -2^(NW-s)) */
               const uint32_t sh = op.kparam;
               const limb_t fm = (((limb_t) 1 << sh) - 1);
               limb_t lowf[NTS];
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane, nx = (i + 1 == NCH) ? 0u : i + 1;
                  ld2(a0[tj], a1[tj], A + 2 * i); ca[tj] = cwA[i];
                  lowf[tj] = A[2 * nx] & fm;
               }
               op_sync<W>(warp);
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  const int64_t r = (int64_t)((limb_t)(int64_t) ca[tj] & fm);             /* c = q 2^s + r, 0 <= r < 2^s */
                  const int64_t e = r + ((i + 1 == NCH) ? -(int64_t) lowf[tj] : (int64_t) lowf[tj]);
                  const limb_t h0 = (a0[tj] >> sh) | (a1[tj] << (64 - sh));
                  const limb_t h1 = (a1[tj] >> sh) | (((limb_t) e & fm) << (64 - sh));
                  st2(S + 2 * i, h0, h1);
                  cwS[i] = (ca[tj] >> sh) + (int32_t)(e >> sh);
               }
            } else
            {  /* MFFT_K_ADD */
#pragma unroll
               for (int tj = 0; tj < NTS; tj++)
               {
                  const uint32_t i = (half + W * tj) * 32u + lane;
                  limb_t x0, x1, y0, y1, r0, r1;
                  ld2(x0, x1, A + 2 * i); ld2(y0, y1, B + 2 * i);
                  const int32_t k = add2(r0, r1, x0, x1, y0, y1) + cwA[i] + cwB[i];
                  st2(S + 2 * i, r0, r1); cwS[i] = k;
               }
            }
            continue;
         }
         uint32_t eSA = op.eSA, eSB = op.eSB, eTA = op.eTA, eTB = op.eTB;
         if (op.cSA | op.cSB | op.cTA | op.cTB)
         {  /* the MFA twist z^(r*c): only the twisted layer pays for the modulo */
            eSA = (eSA + b.col * op.cSA) % M2; eSB = (eSB + b.col * op.cSB) % M2;
            eTA = (eTA + b.col * op.cTA) % M2; eTB = (eTB + b.col * op.cTB) % M2;
         }
         if (!hasB && !hasT && op.sSA != 0)
         {  /* one operand, any rotation (the twist layer): chunk-local */
            uint32_t e = eSA, ng = (op.sSA < 0) ? 1u : 0u;
            if (e >= NW) { e -= NW; ng ^= 1u; }
            rotg_unit<NT, W>(S, A, e, ng, lane, warp);
            continue;
         }
         tterm sa, sb, ta, tb;
         tterm_setup(sa, op.sSA, eSA, NW);
         tterm_setup(sb, hasB ? op.sSB : 0, eSB, NW);
         tterm_setup(ta, hasT ? op.sTA : 0, eTA, NW);
         tterm_setup(tb, (hasT && hasB) ? op.sTB : 0, eTB, NW);
         /* chunk-aligned rotations everywhere, A in every output, and the same A-to-B chunk offset
            in both outputs: each lane works on source chunk pairs */
         uint32_t misal = 0;
         if (sa.present) misal |= eSA;
         if (sb.present) misal |= eSB;
         if (ta.present) misal |= eTA;
         if (tb.present) misal |= eTB;
         bool fast = !(misal & AMASK) && sa.present && (!hasT || ta.present);
         if (fast && sb.present && tb.present && ((sa.yc + tb.yc + 2 * NCH - sb.yc - ta.yc) % NCH) != 0) fast = false;

         if (fast)
         {
            const tterm &ref = sb.present ? sa : ta, &refb = sb.present ? sb : tb;
            const uint32_t dB = (ref.yc + NCH - refb.yc) % NCH;       /* B's chunk = A's chunk + dB */
            const bool useB = sb.present || tb.present;
            const int32_t *cwA = reinterpret_cast<const int32_t *>(A + tile_cfg<NT>::CW);
            const int32_t *cwB = reinterpret_cast<const int32_t *>(B + tile_cfg<NT>::CW);
            limb_t xa[NTS][2], xb[NTS][2]; int32_t ca[NTS], cb[NTS];
#pragma unroll
            for (int tj = 0; tj < NTS; tj++)
            {
               const uint32_t ia = (half + W * tj) * 32u + lane;
#ifdef MFFT_EMU
               xa[tj][0] = A[2 * ia]; xa[tj][1] = A[2 * ia + 1];
#else
               { const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(A + 2 * ia); xa[tj][0] = v.x; xa[tj][1] = v.y; }
#endif
               ca[tj] = cwA[ia];
               xb[tj][0] = 0; xb[tj][1] = 0; cb[tj] = 0;
               if (useB)
               {
                  uint32_t jb = ia + dB; if (jb >= NCH) jb -= NCH;
#ifdef MFFT_EMU
                  xb[tj][0] = B[2 * jb]; xb[tj][1] = B[2 * jb + 1];
#else
                  { const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(B + 2 * jb); xb[tj][0] = v.x; xb[tj][1] = v.y; }
#endif
                  cb[tj] = cwB[jb];
               }
            }
            op_sync<W>(warp);                   /* every lane has read its operands: in-place stores are safe */
#pragma unroll
            for (int tj = 0; tj < NTS; tj++)
            {
               const uint32_t ia = (half + W * tj) * 32u + lane;
               uint32_t jb = ia + dB; if (jb >= NCH) jb -= NCH;
               aligned_out<NT>(S, sa, sb, ia, jb, xa[tj][0], xa[tj][1], ca[tj], xb[tj][0], xb[tj][1], cb[tj]);
               if (hasT) aligned_out<NT>(Tt, ta, tb, ia, jb, xa[tj][0], xa[tj][1], ca[tj], xb[tj][0], xb[tj][1], cb[tj]);
            }
         } else
         {
            limb_t rs[NTS][2], rt[NTS][2]; int32_t ks[NTS], kt[NTS];
#pragma unroll
            for (int tj = 0; tj < NTS; tj++)
            {
               const uint32_t ch = (half + W * tj) * 32u + lane;
               ks[tj] = general_out<NT>(rs[tj][0], rs[tj][1], sa, A, sb, B, ch);
               if (hasT) kt[tj] = general_out<NT>(rt[tj][0], rt[tj][1], ta, A, tb, B, ch);
            }
            op_sync<W>(warp);
            int32_t *cwS = reinterpret_cast<int32_t *>(S + tile_cfg<NT>::CW), *cwT = reinterpret_cast<int32_t *>(Tt + tile_cfg<NT>::CW);
#pragma unroll
            for (int tj = 0; tj < NTS; tj++)
            {
               const uint32_t ch = (half + W * tj) * 32u + lane;
               S[2 * ch] = rs[tj][0]; S[2 * ch + 1] = rs[tj][1]; cwS[ch] = ks[tj];
               if (hasT) { Tt[2 * ch] = rt[tj][0]; Tt[2 * ch + 1] = rt[tj][1]; cwT[ch] = kt[tj]; }
            }
         }
      }
      tile_sync(bar_id, bar_threads);
   }
}

/* store what was written: in place, or gathered (and normalised) into dst */
template <int NT>
__device__ __forceinline__ void tile_store(limb_t *coef, const uint32_t *spos, uint32_t npos, limb_t *slab, const mfft_geom &g,
                                           const mfft_batch &b, uint32_t bi, limb_t *dst, const uint32_t *__restrict__ dstpos,
                                           const uint32_t *__restrict__ dst_base, uint32_t dst_stride, int normalise, bool al16,
                                           uint32_t warp, uint32_t nwarps, uint32_t lane)
{
   constexpr uint32_t L = tile_cfg<NT>::L, SP = tile_cfg<NT>::SP;
   for (uint32_t p = warp; p < npos; p += nwarps)
   {
      const uint32_t pp = spos[p];
      if (!(pp & MFFT_TILE_STORE)) continue;
      limb_t *out;
      if (dst)
      {
         const uint32_t dp = dstpos[pp & MFFT_TILE_POSMASK];
         if (dp == MFFT_NONE) continue;
         out = dst + ((uint64_t) dst_base[bi] + (uint64_t) dp * dst_stride) * g.pitch;
      } else out = tile_block_ptr(slab, g, pp & MFFT_TILE_POSMASK, b);
      limb_t *sblk = coef + (size_t) p * SP;
      if (al16)
      {  /* resolved limbs go straight from registers to HBM, 512 contiguous bytes per warp store */
         limb_t r0[NT], r1[NT];
         int64_t tp = resolve_regs<NT>(r0, r1, sblk, lane);
         if (normalise) tp = normalise_regs<NT>(r0, r1, tp, lane);
#pragma unroll
         for (int ti = 0; ti < NT; ti++) st2(out + 2 * (ti * 32u + lane), r0[ti], r1[ti]);
         if (lane == 0) out[L] = (limb_t) tp;
         continue;
      }
      int64_t top = resolve_carries<NT>(sblk, lane);
      if (normalise) top = normalise_tile(sblk, L, top, lane);
#pragma unroll 4
      for (uint32_t k = lane; k < L; k += 32) out[k] = sblk[k];
      if (lane == 0) out[L] = (limb_t) top;
   }
}

/* tile descriptors of a small pass as kernel parameters (constant bank): no dependent global loads
 * before a CTA can start fetching its coefficients */
#define TP_MAXT 16
#define TP_MAXP 32
#define TP_MAXS 12
#define TP_MAXB 256
struct tile_params {
   uint32_t valid, batch_valid;
   uint32_t debug;          /* developer aid (MPIRFFT_TILE_DEBUG): 1 = skip the stages, 2 = skip the stores (timing floors; wrong results) */
   /* split fused into the tile load (FFT_split_bits, mul_fft.c:115-170): block k of the slab is
      coefficient k = bits [k*bits, (k+1)*bits) of {src, nlimbs}, zero beyond ncoef */
   uint32_t split; const limb_t *split_src; uint64_t split_nlimbs, split_bits, split_ncoef;
   mfft_batch batch[TP_MAXB];
   mfft_tile tiles[TP_MAXT];
   uint32_t pos[TP_MAXT * TP_MAXP];
   uint32_t stoff[TP_MAXT * TP_MAXS];
};

/* what the threads of a tile do while (or instead of) the bulk copies: the split-fused load, the plain
   load of slabs that are not 16-byte aligned, and the zeroing of the carry words */
template <int NT>
__device__ __forceinline__ void tile_fill(limb_t *coef, const uint32_t *spos, uint32_t npos, limb_t *slab, const mfft_geom &g,
                                          const mfft_batch &b, const tile_params &TP, bool bulk_tile,
                                          uint32_t tid, uint32_t nthreads, uint32_t warp, uint32_t nwarps, uint32_t lane,
                                          uint32_t bar_id, uint32_t bar_threads)
{
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH, SP = tile_cfg<NT>::SP;
   if (TP.split)
   {  /* the tile's coefficients are cut straight out of the operand: no split kernel, no slab read */
      /* limb k of coefficient i = bits [i*bits + 64k, +64) of the operand: two loads and a funnel
         shift with the same shift count for the whole coefficient; consecutive threads read
         consecutive limbs, four coefficients are in flight per thread */
      const uint32_t blimbs = (uint32_t)((TP.split_bits + 63) >> 6);     /* limbs that receive bits */
      const limb_t lastmask = (TP.split_bits & 63) ? (((limb_t) 1 << (TP.split_bits & 63)) - 1) : ~(limb_t) 0;
      if (2 * blimbs + 1 <= SP)
      {  /* staged variant: the blimbs+1 raw operand limbs of every coefficient go by 8-byte asynchronous
            copies into the upper end of the coefficient's own buffer (nobody waits on a register for
            them), then they are shifted into place from shared memory, then the rest is zeroed */
         const uint32_t raw = blimbs + 1, STG = SP - raw;
         for (uint32_t t = tid; t < npos * raw; t += nthreads)
         {
            const uint32_t p = t / raw, k = t % raw, pp = spos[p];
            if (!(pp & MFFT_TILE_LOAD)) continue;
            const uint64_t i = (uint64_t) b.base + (uint64_t)(pp & MFFT_TILE_POSMASK) * g.slot_stride;
            const uint64_t q = ((i * TP.split_bits) >> 6) + k;
            limb_t *d = coef + (size_t) p * SP + STG + k;
            if (i < TP.split_ncoef && q < TP.split_nlimbs) cp_async8(d, TP.split_src + q); else *d = 0;
         }
         cp_async_wait_all();
         tile_sync(bar_id, bar_threads);
         for (uint32_t t = tid; t < npos * blimbs; t += nthreads)
         {
            const uint32_t p = t / blimbs, k = t % blimbs, pp = spos[p];
            if (!(pp & MFFT_TILE_LOAD)) continue;
            const uint64_t i = (uint64_t) b.base + (uint64_t)(pp & MFFT_TILE_POSMASK) * g.slot_stride;
            const uint32_t r = (uint32_t)((i * TP.split_bits) & 63);
            const limb_t *st = coef + (size_t) p * SP + STG;
            limb_t v = st[k] >> r;
            if (r) v |= st[k + 1] << (64 - r);
            if (k + 1 == blimbs) v &= lastmask;
            coef[(size_t) p * SP + k] = v;
         }
         tile_sync(bar_id, bar_threads);
         for (uint32_t t = tid; t < npos * (SP - blimbs); t += nthreads)
         {
            const uint32_t p = t / (SP - blimbs), k = blimbs + t % (SP - blimbs);
            if (spos[p] & MFFT_TILE_LOAD) coef[(size_t) p * SP + k] = 0;
         }
      } else
      {
      for (uint32_t p0 = 0; p0 < npos; p0 += 4)
      {
         uint64_t qb[4]; uint32_t rr[4]; bool ld[4], nz[4];
#pragma unroll
         for (int u = 0; u < 4; u++)
         {
            const uint32_t p = p0 + u;
            const uint32_t pp = (p < npos) ? spos[p] : 0u;
            ld[u] = (pp & MFFT_TILE_LOAD) != 0;
            const uint64_t i = (uint64_t) b.base + (uint64_t)(pp & MFFT_TILE_POSMASK) * g.slot_stride;
            const uint64_t off = i * TP.split_bits;
            nz[u] = ld[u] && i < TP.split_ncoef;
            qb[u] = off >> 6; rr[u] = (uint32_t)(off & 63);
         }
         for (uint32_t k = tid; k < L; k += nthreads)
         {
            limb_t lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
               lo[u] = 0; hi[u] = 0;
               if (nz[u] && k < blimbs)
               {
                  const uint64_t q = qb[u] + k;
                  if (q < TP.split_nlimbs) lo[u] = TP.split_src[q];
                  if (q + 1 < TP.split_nlimbs) hi[u] = TP.split_src[q + 1];
               }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
               if (!ld[u]) continue;
               limb_t v = lo[u] >> rr[u];
               if (rr[u]) v |= hi[u] << (64 - rr[u]);
               if (k + 1 == blimbs) v &= lastmask;
               coef[(size_t)(p0 + u) * SP + k] = v;
            }
         }
      }
      for (uint32_t t = tid; t < npos * NCH; t += nthreads)
         if (spos[t / NCH] & MFFT_TILE_LOAD) reinterpret_cast<int32_t *>(coef + (size_t)(t / NCH) * SP + tile_cfg<NT>::CW)[t % NCH] = 0;
      }
   } else
   if (bulk_tile)
   {  /* the block images are on their way; meanwhile zero the carry words (they live behind the image) */
      for (uint32_t t = tid; t < npos * (NCH / 4); t += nthreads)
      {
         const uint32_t p = t / (NCH / 4), c = t % (NCH / 4);
         if (!(spos[p] & MFFT_TILE_LOAD)) continue;
         st2(coef + (size_t) p * SP + tile_cfg<NT>::CW + 2 * c, 0, 0);
      }
   } else
   for (uint32_t p = warp; p < npos; p += nwarps)
   {
      const uint32_t pp = spos[p];
      if (!(pp & MFFT_TILE_LOAD)) continue;
      const limb_t *src = tile_block_ptr(slab, g, pp & MFFT_TILE_POSMASK, b);
      limb_t *d = coef + (size_t) p * SP;
#pragma unroll 4
      for (uint32_t k = lane; k < L; k += 32) d[k] = src[k];
      int32_t *cw = reinterpret_cast<int32_t *>(d + tile_cfg<NT>::CW);
#pragma unroll
      for (uint32_t ch = lane; ch < NCH; ch += 32) cw[ch] = (ch == NCH - 1) ? (int32_t)(int64_t) src[L] : 0;
   }
}

/* issued by one warp: arm the barrier with the bytes of the tile's op descriptors and block images,
   then fire the bulk copies (one per coefficient that is read before it is written) */
template <int NT, class POS>
__device__ __forceinline__ void tile_issue(tile_mbar *mbar, mfft_tileop *sops, limb_t *coef, const mfft_tile &T, POS posf,
                                           const mfft_tileop *__restrict__ ops, limb_t *slab, const mfft_geom &g,
                                           const mfft_batch &b, bool bulk_tile, uint32_t lane)
{
   constexpr uint32_t L = tile_cfg<NT>::L, SP = tile_cfg<NT>::SP;
   uint32_t nload = 0;
   if (bulk_tile) for (uint32_t p = lane; p < T.npos; p += 32) nload += (posf(p) & MFFT_TILE_LOAD) ? 1u : 0u;
#pragma unroll
   for (int off = 16; off; off >>= 1) nload += __shfl_xor_sync(FULL, nload, off);
   if (lane == 0)
   {
      mbar_arrive_expect_tx(mbar, T.nops * (uint32_t) sizeof(mfft_tileop) + nload * (L + 2) * 8u);
      if (T.nops) bulk_g2s(sops, ops + T.op_off, T.nops * (uint32_t) sizeof(mfft_tileop), mbar);
   }
   __syncwarp();
   if (bulk_tile)
      for (uint32_t p = lane; p < T.npos; p += 32)
      {
         const uint32_t pp = posf(p);
         if (pp & MFFT_TILE_LOAD) bulk_g2s(coef + (size_t) p * SP, tile_block_ptr(slab, g, pp & MFFT_TILE_POSMASK, b), (L + 2) * 8u, mbar);
      }
}

/* ---- the kernel ----------------------------------------------------------------------------- */
/* WSPLIT = 2: two warps share every op (each takes every other group of 32 chunks), so a tile is worked
   on by twice as many warps holding half as many chunks each: 8 warps x 64 registers per 16-coefficient
   tile, four CTAs per SM = 32 resident warps instead of 16 */
template <int NT, int NTHREADS, int WSPLIT>
__global__ void __launch_bounds__(NTHREADS, (NTHREADS == 128 || WSPLIT == 2) ? 4 : 2)
k_run_tiles(limb_t *slab, mfft_geom g, const mfft_tile *__restrict__ tiles, const uint32_t *__restrict__ pos,
            const mfft_tileop *__restrict__ ops, const mfft_batch *__restrict__ batch, uint32_t nbatch,
            limb_t *dst, const uint32_t *__restrict__ dstpos, const uint32_t *__restrict__ dst_base,
            uint32_t dst_stride, int normalise, uint32_t desc_bytes, const uint32_t *__restrict__ stoff,
            unsigned long long *timing, const __grid_constant__ tile_params TP)
{
#ifndef MFFT_EMU
#define TILE_STAMP(k) do { if (timing && threadIdx.x == 0) { unsigned long long t__; \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); timing[(size_t) blockIdx.x * 8 + (k)] = t__; } } while (0)
#else
#define TILE_STAMP(k) do { } while (0)
#endif
   TILE_STAMP(0);
   MFFT_DYN_SMEM(limb_t, sm);
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH, SP = tile_cfg<NT>::SP;
   constexpr uint32_t NW = 64u * L, M2 = 2u * NW;
   constexpr uint32_t AMASK = 127u;                           /* exponent bits that break chunk alignment */
   const uint32_t bi = blockIdx.x % nbatch, tix = blockIdx.x / nbatch;
   const mfft_tile T = TP.valid ? TP.tiles[tix] : tiles[tix];
   const mfft_batch b = TP.batch_valid ? TP.batch[bi] : batch[bi];
   const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
   /* shared memory: [mbarrier] [op descriptors | position list | stage offsets] [coefficients] */
   tile_mbar *mbar = (tile_mbar *) sm;
   mfft_tileop *sops = (mfft_tileop *)(sm + MBAR_BYTES / 8);
   uint32_t *spos = (uint32_t *)(sops + T.nops);
   uint32_t *sst = spos + T.npos;                      /* [nstages+1] first op of each stage */
   limb_t *coef = sm + desc_bytes / 8;

   {  /* the barrier all of the tile's bulk copies complete on; position list and stage offsets */
      static_assert(sizeof(mfft_tileop) % 16 == 0, "op descriptors are fetched by one bulk copy");
      if (tid == 0) mbar_init(mbar, 1);
      if (TP.valid)
      {
         for (uint32_t k = tid; k < T.npos; k += blockDim.x) spos[k] = TP.pos[tix * TP_MAXP + k];
         for (uint32_t k = tid; k <= T.nstages; k += blockDim.x) sst[k] = TP.stoff[tix * TP_MAXS + k];
      } else
      {
         for (uint32_t k = tid; k < T.npos; k += blockDim.x) spos[k] = pos[T.pos_off + k];
         for (uint32_t k = tid; k <= T.nstages; k += blockDim.x) sst[k] = stoff[T.pad + k];
      }
   }
   __syncthreads();                                   /* the position list is read by everybody below */
#ifndef MFFT_EMU
   /* programmatic dependent launch: when the launch carries the stream-serialisation attribute this
      CTA may have been scheduled (and its descriptors staged) while the previous pass was still
      draining; nothing written by that pass is read above.  Without the attribute both are no-ops. */
   asm volatile("griddepcontrol.wait;" ::: "memory");
   asm volatile("griddepcontrol.launch_dependents;");
#endif
   /* load the positions that are read before being written: the body as it is, carry words 0
      except the last one, which is the block's signed top limb */
   const bool al16 = ((g.pitch & 1u) == 0) && ((reinterpret_cast<uintptr_t>(slab) & 15u) == 0) &&
                     (dst == nullptr || (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
   const bool bulk_tile = al16 && !TP.split && g.pitch >= L + 2;      /* whole block images by bulk copy */
   if (warp == 0)     /* one warp issues the tile's copies */
      tile_issue<NT>(mbar, sops, coef, T, [&](uint32_t p) { return spos[p]; }, ops, slab, g, b, bulk_tile, lane);
   tile_fill<NT>(coef, spos, T.npos, slab, g, b, TP, bulk_tile, tid, blockDim.x, warp, nwarps, lane, 0u, 0u);
   TILE_STAMP(1);
   /* one warp polls the mbarrier; the others sleep in the hardware barrier below instead of spending issue
      slots on try_wait loops next to CTAs that are computing (TP.debug & 4: every warp polls, for A/B) */
   if (warp == 0 || (TP.debug & 4u)) mbar_wait(mbar, 0);
   if (bulk_tile)
   {  /* last carry word = the block's signed top limb, which arrived with the image */
      __syncthreads();
      if (tid < T.npos && (spos[tid] & MFFT_TILE_LOAD))
         reinterpret_cast<int32_t *>(coef + (size_t) tid * SP + tile_cfg<NT>::CW)[NCH - 1] = (int32_t)(int64_t) coef[(size_t) tid * SP + L];
   }
   __syncthreads();
   TILE_STAMP(2);

   if (!(TP.debug & 1u)) tile_stages<NT, WSPLIT>(coef, sops, sst, T.nstages, b, warp, nwarps, lane, 0u, 0u);

   TILE_STAMP(3);
   /* store what was written: in place, or gathered (and normalised) into dst */
   if (!(TP.debug & 2u)) tile_store<NT>(coef, spos, T.npos, slab, g, b, bi, dst, dstpos, dst_base, dst_stride, normalise, al16, warp, nwarps, lane);
#ifndef MFFT_EMU
   if (timing)
   {
      __syncthreads();
      TILE_STAMP(4);
      if (threadIdx.x == 0) { uint32_t sm__; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm__)); timing[(size_t) blockIdx.x * 8 + 5] = sm__; }
   }
#endif
#undef TILE_STAMP
}


/* ---- the persistent, pipelined variant ------------------------------------------------------- */
/* One CTA per SM walks over tiles w = blockIdx.x, blockIdx.x + gridDim.x, ...  The CTA's shared memory
 * is a ring of TPP_NBUF tile buffers; TPP_GROUPS groups of four warps each work on one tile at a time
 * (item k of the CTA goes to group k mod TPP_GROUPS and lives in buffer k mod TPP_NBUF), so while the
 * groups run their stages and stores, the spare buffers are being filled by the copy engine: the
 * group that finishes item k hands its buffer straight back by issuing the bulk copies of item
 * k + TPP_NBUF.  Without this, all CTAs of a launch move through load -> stages -> store in lockstep
 * waves (measured on B200 at 2^20 limbs: load 6 us + stages 12 us + stores 6 us per pass add up instead
 * of overlapping).  full[b]: mbarrier the copies of buffer b complete on; armed[b]: how many uses of
 * buffer b have been armed -- a consumer first waits for its use to be armed, then for the phase
 * (groups consume different buffers, so a phase parity alone could be one lap behind). */
#define TPP_GROUPS 4
#define TPP_NBUF 5
#define TPP_CTL_BYTES 256

template <int NT>
__global__ void __launch_bounds__(TPP_GROUPS * 128, 1)
k_run_tiles_p(limb_t *slab, mfft_geom g, const mfft_tile *__restrict__ tiles, const uint32_t *__restrict__ pos,
              const mfft_tileop *__restrict__ ops, const mfft_batch *__restrict__ batch, uint32_t nbatch,
              limb_t *dst, const uint32_t *__restrict__ dstpos, const uint32_t *__restrict__ dst_base,
              uint32_t dst_stride, int normalise, uint32_t desc_bytes, const uint32_t *__restrict__ stoff,
              uint32_t total, uint32_t buf_bytes, const __grid_constant__ tile_params TP)
{
   MFFT_DYN_SMEM(limb_t, sm);
   constexpr uint32_t L = tile_cfg<NT>::L, NCH = tile_cfg<NT>::NCH, SP = tile_cfg<NT>::SP;
   unsigned char *smb = (unsigned char *) sm;
   const uint32_t tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127u, lane = tid & 31u, gwarp = gtid >> 5;
   const uint32_t K = (total > blockIdx.x) ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   /* my items */
   volatile uint32_t *armed = (volatile uint32_t *)(smb + TPP_NBUF * MBAR_BYTES);
   static_assert(TPP_NBUF * MBAR_BYTES + TPP_NBUF * 4 <= TPP_CTL_BYTES, "control block");
   if (tid == 0)
      for (uint32_t bf = 0; bf < TPP_NBUF; bf++) { mbar_init((tile_mbar *)(smb + bf * MBAR_BYTES), 1); armed[bf] = 0; }
   __syncthreads();
#ifndef MFFT_EMU
   asm volatile("griddepcontrol.wait;" ::: "memory");           /* see k_run_tiles */
   asm volatile("griddepcontrol.launch_dependents;");
#endif
   const bool al16 = ((g.pitch & 1u) == 0) && ((reinterpret_cast<uintptr_t>(slab) & 15u) == 0) &&
                     (dst == nullptr || (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
   const bool bulk_tile = al16 && !TP.split && g.pitch >= L + 2;

   /* arm and fire the copies of item k (one warp) */
   auto issue = [&](uint32_t k)
   {
      const uint32_t w = blockIdx.x + k * gridDim.x, bi = w % nbatch, tix = w / nbatch, bf = k % TPP_NBUF;
      const mfft_tile T = TP.valid ? TP.tiles[tix] : tiles[tix];
      const mfft_batch b = TP.batch_valid ? TP.batch[bi] : batch[bi];
      unsigned char *bb = smb + TPP_CTL_BYTES + (size_t) bf * buf_bytes;
      tile_mbar *mbar = (tile_mbar *)(smb + bf * MBAR_BYTES);
      fence_async_smem();                              /* the buffer was last touched through the generic proxy */
      if (TP.valid) tile_issue<NT>(mbar, (mfft_tileop *) bb, (limb_t *)(bb + desc_bytes), T, [&](uint32_t p) { return TP.pos[tix * TP_MAXP + p]; }, ops, slab, g, b, bulk_tile, lane);
      else          tile_issue<NT>(mbar, (mfft_tileop *) bb, (limb_t *)(bb + desc_bytes), T, [&](uint32_t p) { return pos[T.pos_off + p]; }, ops, slab, g, b, bulk_tile, lane);
      if (lane == 0) { __threadfence_block(); armed[bf] = k / TPP_NBUF + 1; }
   };
   if (gwarp == 0)
      for (uint32_t k = grp; k < K && k < TPP_NBUF; k += TPP_GROUPS) issue(k);

   for (uint32_t k = grp; k < K; k += TPP_GROUPS)
   {
      const uint32_t w = blockIdx.x + k * gridDim.x, bi = w % nbatch, tix = w / nbatch, bf = k % TPP_NBUF, use = k / TPP_NBUF;
      const mfft_tile T = TP.valid ? TP.tiles[tix] : tiles[tix];
      const mfft_batch b = TP.batch_valid ? TP.batch[bi] : batch[bi];
      unsigned char *bb = smb + TPP_CTL_BYTES + (size_t) bf * buf_bytes;
      tile_mbar *mbar = (tile_mbar *)(smb + bf * MBAR_BYTES);
      mfft_tileop *sops = (mfft_tileop *) bb;
      uint32_t *spos = (uint32_t *)(sops + T.nops);
      uint32_t *sst = spos + T.npos;
      limb_t *coef = (limb_t *)(bb + desc_bytes);
      /* the buffer is mine once the copies of my item have been armed (its previous user did that
         after its last access) */
      while (armed[bf] < use + 1)
      {
#ifdef MFFT_EMU
         emu_yield();
#else
         __nanosleep(64);
#endif
      }
      __threadfence_block();
      if (TP.valid)
      {
         for (uint32_t q = gtid; q < T.npos; q += 128) spos[q] = TP.pos[tix * TP_MAXP + q];
         for (uint32_t q = gtid; q <= T.nstages; q += 128) sst[q] = TP.stoff[tix * TP_MAXS + q];
      } else
      {
         for (uint32_t q = gtid; q < T.npos; q += 128) spos[q] = pos[T.pos_off + q];
         for (uint32_t q = gtid; q <= T.nstages; q += 128) sst[q] = stoff[T.pad + q];
      }
      tile_sync(1 + grp, 128);
      tile_fill<NT>(coef, spos, T.npos, slab, g, b, TP, bulk_tile, gtid, 128, gwarp, 4, lane, 1 + grp, 128);
      mbar_wait(mbar, use & 1u);
      if (bulk_tile)
      {  /* last carry word = the block's signed top limb, which arrived with the image */
         tile_sync(1 + grp, 128);
         if (gtid < T.npos && (spos[gtid] & MFFT_TILE_LOAD))
            reinterpret_cast<int32_t *>(coef + (size_t) gtid * SP + tile_cfg<NT>::CW)[NCH - 1] = (int32_t)(int64_t) coef[(size_t) gtid * SP + L];
      }
      tile_sync(1 + grp, 128);
      if (!(TP.debug & 1u)) tile_stages<NT, 1>(coef, sops, sst, T.nstages, b, gwarp, 4, lane, 1 + grp, 128);
      if (!(TP.debug & 2u)) tile_store<NT>(coef, spos, T.npos, slab, g, b, bi, dst, dstpos, dst_base, dst_stride, normalise, al16, gwarp, 4, lane);
      tile_sync(1 + grp, 128);                         /* nobody of the group reads the buffer any more */
      if (gwarp == 0 && k + TPP_NBUF < K) issue(k + TPP_NBUF);
   }
}

#endif
