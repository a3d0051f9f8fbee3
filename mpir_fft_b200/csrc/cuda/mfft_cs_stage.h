/* mfft_cs_stage.h -- carry-save radix-2 layers on HBM-resident coefficients of any size
 * (included by mfft_kernels.cu after mfft_tiles.h, whose chunk arithmetic it shares).
 *
 * Coefficient rings above 512 limbs (products beyond 1.6e7 limbs: the 2^24 .. 2^30-limb range of
 * BASELINE configs[4]) do not fit the shared-memory tiles; every radix-2 layer is then one launch
 * that reads and writes the live slab once.  The layer kernel keeps the coefficients in the same
 * carry-save form as the tile executor -- 128-bit chunks, one signed 32-bit carry word per chunk
 * in a side array -- so that a butterfly is chunk-local: thread i adds / subtracts chunk i of the
 * two operands and stores the results at (rotated) chunk positions of the two outputs.  The slot
 * view of the schedules writes every output into the other half of the slab (sched.c: emit), so
 * no chunk is read after it has been overwritten.  Carries are resolved once per transform, in the
 * gather (k_finalize_cs), which also normalises (mpn_normmod_2expp1, mul_fft.c:272-294).
 */
#ifndef MFFT_CS_STAGE_H
#define MFFT_CS_STAGE_H

__device__ __forceinline__ uint64_t cs_block_index(const mfft_geom &g, uint32_t slot, const mfft_batch &b)
{
   const uint32_t half = (slot / g.S) ^ b.parity;
   return (uint64_t) half * g.half_blocks + b.base + (uint64_t)(slot % g.S) * g.slot_stride;
}

struct cs_blk { limb_t *x; int32_t *c; };
__device__ __forceinline__ cs_blk cs_block(limb_t *slab, int32_t *cw, const mfft_geom &g, uint32_t slot, const mfft_batch &b)
{
   const uint64_t idx = cs_block_index(g, slot, b);
   cs_blk r; r.x = slab + idx * g.pitch; r.c = cw + idx * (g.l / 2);
   return r;
}
__device__ __forceinline__ cval csb_load(const cs_blk &B, uint32_t ch)
{ cval r; ld2(r.x0, r.x1, B.x + 2 * ch); r.c = B.c[ch]; return r; }
__device__ __forceinline__ void csb_store(const cs_blk &B, uint32_t ch, const cval &v)
{ st2(B.x + 2 * ch, v.x0, v.x1); B.c[ch] = v.c; }
__device__ __forceinline__ cval cv_neg(const cval &a)
{ cval r; const int32_t bw = sub2(r.x0, r.x1, 0, 0, a.x0, a.x1); r.c = bw - a.c; return r; }

/* carry words of the blocks a transform starts from: zero, except the last one = the block's top limb */
__global__ void __launch_bounds__(256)
k_cs_init(const limb_t *slab, int32_t *cw, mfft_geom g, const mfft_batch *__restrict__ batch, uint32_t nbatch)
{
   const uint32_t NCH = g.l / 2;
   const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
   const uint64_t total = (uint64_t) g.S * nbatch * NCH;
   if (t >= total) return;
   const uint32_t ch = (uint32_t)(t % NCH);
   const uint64_t pb = t / NCH;
   const mfft_batch b = batch[pb % nbatch];
   const uint64_t idx = cs_block_index(g, (uint32_t)(pb / nbatch), b);
   cw[idx * NCH + ch] = (ch == NCH - 1) ? (int32_t)(int64_t) slab[idx * g.pitch + g.l] : 0;
}

/* one op on carry-save blocks: reads A (and B), writes S (and T).  All reads go through the cs_blk
   handles, so the operands may be shared-memory copies (in-place execution, k_stage_cs_ip) */
__device__ __forceinline__ void cs_op_apply(const mfft_op &op, const cs_blk &A, const cs_blk &B, const cs_blk &S, const cs_blk &T,
                                            const mfft_batch &b, const uint32_t NCH, const uint32_t NW)
{
   const uint32_t yc = op.kparam & 0x7fffffffu, neg = op.kparam >> 31;
   switch (op.kind)
   {
   case MFFT_K_FWD:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const cval a = csb_load(A, i), bb = csb_load(B, i);
         csb_store(S, i, cv_add(a, bb));
         uint32_t o = i + yc, n = neg;
         if (o >= NCH) { o -= NCH; n ^= 1u; }
         csb_store(T, o, n ? cv_sub(bb, a) : cv_sub(a, bb));
      }
      break;
   case MFFT_K_INV:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const uint32_t j = (i >= yc) ? i - yc : i + NCH - yc;
         const cval a = csb_load(A, i), bb = csb_load(B, j);
         const bool n = ((i < yc) ? 1u : 0u) != neg;
         csb_store(n ? T : S, i, cv_add(a, bb));
         csb_store(n ? S : T, i, cv_sub(a, bb));
      }
      break;
   case MFFT_K_ROT:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const cval a = csb_load(A, i);
         uint32_t o = i + yc, n = neg;
         if (o >= NCH) { o -= NCH; n ^= 1u; }
         csb_store(S, o, n ? cv_neg(a) : a);
      }
      break;
   case MFFT_K_ADD:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x) csb_store(S, i, cv_add(csb_load(A, i), csb_load(B, i)));
      break;
   case MFFT_K_2AMB:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const cval a = csb_load(A, i), bb = csb_load(B, i);
         cval d; d.x0 = a.x0 << 1; d.x1 = (a.x1 << 1) | (a.x0 >> 63); d.c = 2 * a.c + (int32_t)(a.x1 >> 63);
         csb_store(S, i, cv_sub(d, bb));
         if (op.outT != MFFT_NONE)
         {
            uint32_t o = i + yc, n = neg;
            if (o >= NCH) { o -= NCH; n ^= 1u; }
            csb_store(T, o, n ? cv_sub(bb, a) : cv_sub(a, bb));
         }
      }
      break;
   case MFFT_K_DBL:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const cval a = csb_load(A, i); cval r;
         r.x0 = a.x0 << 1; r.x1 = (a.x1 << 1) | (a.x0 >> 63); r.c = 2 * a.c + (int32_t)(a.x1 >> 63);
         csb_store(S, i, r);
      }
      break;
   case MFFT_K_HALF:
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const uint32_t nx = (i + 1 == NCH) ? 0u : i + 1;
         const cval t = cv_add(csb_load(A, i), csb_load(B, i));
         const int32_t lowbit = (int32_t)((A.x[2 * nx] ^ B.x[2 * nx]) & 1u);
         const int32_t e = (t.c & 1) + ((i + 1 == NCH) ? -lowbit : lowbit);
         cval r;
         r.x0 = (t.x0 >> 1) | (t.x1 << 63); r.x1 = (t.x1 >> 1) | ((limb_t)(e & 1) << 63); r.c = (t.c >> 1) + (e >> 1);
         csb_store(S, i, r);
      }
      break;
   case MFFT_K_SHR:
   {
      const uint32_t sh = op.kparam; const limb_t fm = (((limb_t) 1 << sh) - 1);
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         const uint32_t nx = (i + 1 == NCH) ? 0u : i + 1;
         const cval a = csb_load(A, i);
         const limb_t lowf = A.x[2 * nx] & fm;
         const int64_t r0 = (int64_t)((limb_t)(int64_t) a.c & fm);
         const int64_t e = r0 + ((i + 1 == NCH) ? -(int64_t) lowf : (int64_t) lowf);
         cval r;
         r.x0 = (a.x0 >> sh) | (a.x1 << (64 - sh)); r.x1 = (a.x1 >> sh) | (((limb_t) e & fm) << (64 - sh));
         r.c = (a.c >> sh) + (int32_t)(e >> sh);
         csb_store(S, i, r);
      }
      break;
   }
   default:     /* MFFT_K_ANY with one operand: rotation by any number of bits (see rotg_unit) */
   {
      uint32_t e = (uint32_t)((op.eSA + (uint64_t) b.col * op.cSA) % (2ull * NW)), ng = (op.sSA < 0) ? 1u : 0u;
      if (e >= NW) { e -= NW; ng ^= 1u; }
      uint32_t yc1 = (e + 127u) >> 7;
      const uint32_t s = 128u * yc1 - e;
      if (yc1 == NCH) { yc1 = 0; ng ^= 1u; }
      for (uint32_t j = threadIdx.x; j < NCH; j += blockDim.x)
      {
         const uint32_t wj = (j < yc1) ? 1u : 0u, sj = wj ? j + NCH - yc1 : j - yc1;
         cval r = csb_load(A, sj);
         if (wj != ng) r = cv_neg(r);
         if (s == 0) { csb_store(S, j, r); continue; }
         const uint32_t j1 = (j + 1 == NCH) ? 0u : j + 1;
         const uint32_t w1 = (j1 < yc1) ? 1u : 0u, s1 = w1 ? j1 + NCH - yc1 : j1 - yc1;
         limb_t y0, y1;
         ld2(y0, y1, A.x + 2 * s1);
         if (w1 != ng) (void) sub2(y0, y1, 0, 0, y0, y1);
         limb_t h0, h1, f0, f1, g0, g1; cval o;
         const int64_t c64 = (int64_t) r.c;
         shr128(h0, h1, r.x0, r.x1, s);
         shl128(f0, f1, (limb_t) c64, (limb_t)(c64 >> 63), 128u - s);
         shl128(g0, g1, y0, y1, 128u - s);
         h0 |= f0; h1 |= f1;
         const int32_t k = (j + 1 == NCH) ? sub2(o.x0, o.x1, h0, h1, g0, g1) : add2(o.x0, o.x1, h0, h1, g0, g1);
         o.c = ((s >= 31) ? (r.c >> 31) : (r.c >> s)) + k;
         csb_store(S, j, o);
      }
      break;
   }
   }
}

/* one layer: CTA = (op, batch entry), threads stride over the chunks */
__global__ void __launch_bounds__(256)
k_stage_cs(limb_t *slab, int32_t *cw, mfft_geom g, const mfft_op *__restrict__ ops, uint32_t count,
           const mfft_batch *__restrict__ batch, uint32_t nbatch)
{
   const uint32_t NCH = g.l / 2, NW = 64u * g.l;
   const mfft_op op = ops[blockIdx.x / nbatch];
   const mfft_batch b = batch[blockIdx.x % nbatch];
   const cs_blk A = cs_block(slab, cw, g, op.inA, b);
   const cs_blk B = (op.inB != MFFT_NONE) ? cs_block(slab, cw, g, op.inB, b) : A;
   const cs_blk S = cs_block(slab, cw, g, op.outS, b);
   const cs_blk T = (op.outT != MFFT_NONE) ? cs_block(slab, cw, g, op.outT, b) : S;
   (void) count;
   cs_op_apply(op, A, B, S, T, b, NCH, NW);
}

/* The same IN PLACE on physical positions (pA, pB -> pS, pT of slab half 0): the operands are staged
 * in shared memory first, so an output may overwrite an input.  For the few ops of a big-ring
 * transform that are not slice-local (bit-granular twist rotations, halving, the final scaling);
 * everything else runs in the multi-layer sliced passes (k_run_tiles_sliced).  Dynamic shared memory:
 * two operands of l limbs + l/2 carry words. */
__global__ void __launch_bounds__(1024)
k_stage_cs_ip(limb_t *slab, int32_t *cw, mfft_geom g, const mfft_op *__restrict__ ops, uint32_t count,
              const mfft_batch *__restrict__ batch, uint32_t nbatch)
{
   MFFT_DYN_SMEM(limb_t, sm);
   const uint32_t NCH = g.l / 2, NW = 64u * g.l;
   const mfft_op op = ops[blockIdx.x / nbatch];
   const mfft_batch b = batch[blockIdx.x % nbatch];
   const cs_blk Ag = cs_block(slab, cw, g, op.pA, b);
   const bool hasB = (op.pB != MFFT_NONE) && (op.pB != op.pA);
   const cs_blk Bg = hasB ? cs_block(slab, cw, g, op.pB, b) : Ag;
   const cs_blk S = cs_block(slab, cw, g, op.pS, b);
   const cs_blk T = (op.pT != MFFT_NONE) ? cs_block(slab, cw, g, op.pT, b) : S;
   const uint32_t stride = g.l + g.l / 4;                 /* limbs per staged operand: body + carry words */
   /* only an operand that an output overwrites has to be staged (e.g. the rows a truncated transform
      synthesises from others, 1217-1220, overwrite nothing: no staging at all) */
   const bool stA = (op.pS == op.pA) || (op.pT == op.pA);
   const bool stB = hasB && ((op.pS == op.pB) || (op.pT == op.pB));
   cs_blk A = Ag, B = Bg;
   (void) count;
   if (stA) { A.x = sm; A.c = reinterpret_cast<int32_t *>(sm + g.l); }
   if (stB) { limb_t *q = sm + (stA ? stride : 0); B.x = q; B.c = reinterpret_cast<int32_t *>(q + g.l); }
   if (!hasB) B = A;
   if (stA || stB)
   {
      for (uint32_t i = threadIdx.x; i < NCH; i += blockDim.x)
      {
         limb_t x0, x1;
         if (stA) { ld2(x0, x1, Ag.x + 2 * i); st2(A.x + 2 * i, x0, x1); A.c[i] = Ag.c[i]; }
         if (stB) { ld2(x0, x1, Bg.x + 2 * i); st2(B.x + 2 * i, x0, x1); B.c[i] = Bg.c[i]; }
      }
      __syncthreads();
   }
   cs_op_apply(op, A, B, S, T, b, NCH, NW);
}

/* ---- multi-layer passes on chunk slices of big coefficients -------------------------------------
 * All inner layers of a big-ring transform rotate by whole chunks, and by multiples of gs chunks:
 * then the chunks { i0 + gs t } of every coefficient of a sub-transform form a closed system -- a
 * coefficient of the ring with NCH/gs chunks (rotation by gs q chunks = rotation of the slice by q,
 * the wrap-around negation included).  A CTA takes (tile of positions, batch entry, slice i0), loads
 * those chunks and their carry words into shared memory, runs several layers there with the tile
 * executor's chunk-local kinds, and stores chunks and carry words back in place: one read and one
 * write of the slab per PASS instead of per layer.  (North-star: "several radix-2 layers ... inside
 * one shared-memory tile"; the slice makes that possible when one coefficient is 16-64 KB.) */
template <int NT, int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, 2)
k_run_tiles_sliced(limb_t *slab, int32_t *cw, mfft_geom g, uint32_t gs, uint32_t R, const mfft_tile *__restrict__ tiles,
                   const uint32_t *__restrict__ pos, const mfft_tileop *__restrict__ ops, const uint32_t *__restrict__ stoff,
                   const mfft_batch *__restrict__ batch, uint32_t nbatch, uint32_t desc_bytes)
{
   /* R adjacent slices i0 .. i0+R-1 per CTA: their chunks are contiguous in HBM (R x 16 bytes of body,
      R x 4 bytes of carry words per step of gs chunks), which is what makes the sectors full */
   MFFT_DYN_SMEM(limb_t, sm);
   constexpr uint32_t NCHV = tile_cfg<NT>::NCH, SP = tile_cfg<NT>::SP, CW = tile_cfg<NT>::CW;
   const uint32_t nsl = gs / R;
   const uint32_t i0 = (blockIdx.x % nsl) * R, rest = blockIdx.x / nsl, bi = rest % nbatch, tix = rest / nbatch;
   const mfft_tile T = tiles[tix];
   const mfft_batch b = batch[bi];
   const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NTHREADS >> 5;
   const uint32_t NCH = g.l / 2;
   mfft_tileop *sops = (mfft_tileop *) sm;
   uint32_t *spos = (uint32_t *)(sops + T.nops);
   uint32_t *sst = spos + T.npos;
   limb_t *coef = sm + desc_bytes / 8;               /* slice r of position p: coef + (r * T.npos + p) * SP */
   {
      const limb_t *src = (const limb_t *)(ops + T.op_off);
      limb_t *d = (limb_t *) sops;
      for (uint32_t k = tid; k < T.nops * (uint32_t)(sizeof(mfft_tileop) / 16); k += NTHREADS) cp_async16(d + 2 * k, src + 2 * k);
      for (uint32_t k = tid; k < T.npos; k += NTHREADS) spos[k] = pos[T.pos_off + k];
      for (uint32_t k = tid; k <= T.nstages; k += NTHREADS) sst[k] = stoff[T.pad + k];
   }
   __syncthreads();
   /* a warp per coefficient: the block address is computed once, the lanes walk the slices' chunks
      (16-byte chunk bodies and 4-byte carry words, both by asynchronous copies) */
   for (uint32_t p = warp; p < T.npos; p += nwarps)
   {
      const uint32_t pp = spos[p];
      if (!(pp & MFFT_TILE_LOAD)) continue;
      const uint64_t idx = cs_block_index(g, pp & MFFT_TILE_POSMASK, b);
      const limb_t *src = slab + idx * g.pitch + 2 * (size_t) i0;
      const int32_t *csrc = cw + idx * NCH + i0;
#pragma unroll 4
      for (uint32_t q = lane; q < NCHV * R; q += 32)
      {
         const uint32_t c = q / R, r = q % R;
         limb_t *d = coef + ((size_t) r * T.npos + p) * SP;
         cp_async16(d + 2 * c, src + 2 * ((size_t) gs * c + r));
         cp_async4(reinterpret_cast<int32_t *>(d + CW) + c, csrc + (size_t) gs * c + r);
      }
   }
   cp_async_wait_all();
   __syncthreads();
   for (uint32_t r = 0; r < R; r++)
      tile_stages<NT, 1>(coef + (size_t) r * T.npos * SP, sops, sst, T.nstages, b, warp, nwarps, lane, 0u, 0u);
   for (uint32_t p = warp; p < T.npos; p += nwarps)
   {
      const uint32_t pp = spos[p];
      if (!(pp & MFFT_TILE_STORE)) continue;
      const uint64_t idx = cs_block_index(g, pp & MFFT_TILE_POSMASK, b);
      limb_t *dstp = slab + idx * g.pitch + 2 * (size_t) i0;
      int32_t *cdst = cw + idx * NCH + i0;
#pragma unroll 4
      for (uint32_t q = lane; q < NCHV * R; q += 32)
      {
         const uint32_t c = q / R, r = q % R;
         const limb_t *d = coef + ((size_t) r * T.npos + p) * SP;
         limb_t x0, x1;
         ld2(x0, x1, d + 2 * c);
         st2(dstp + 2 * ((size_t) gs * c + r), x0, x1);
         cdst[(size_t) gs * c + r] = reinterpret_cast<const int32_t *>(d + CW)[c];
      }
   }
}

/* gather + carry resolution + optional normalisation: CTA = (move, batch entry).  Every thread owns
 * a run of consecutive chunks; the carries between the runs are a block-wide prefix scan of the
 * three-state transition functions of ripple_regs (mfft_tiles.h).  Dynamic shared memory: 2 bytes
 * per chunk (leftover carry in {-1,0,1}, all-ones / all-zero flags). */
__device__ __forceinline__ uint32_t cs_tf(int32_t d, uint32_t fl)
{ return (uint32_t)(d - (int32_t)((fl >> 1) & 1u) + 1) | ((uint32_t)(d + 1) << 2) | ((uint32_t)(d + (int32_t)(fl & 1u) + 1) << 4); }

__global__ void __launch_bounds__(256)
k_finalize_cs(limb_t *dst, uint32_t dst_stride, const uint32_t *__restrict__ dst_base, limb_t *slab, int32_t *cw,
              mfft_geom g, const mfft_move *__restrict__ moves, uint32_t nmoves, const mfft_batch *__restrict__ batch,
              uint32_t nbatch, int normalise)
{
   MFFT_DYN_SMEM(limb_t, smraw);
   __shared__ uint32_t sF[256]; __shared__ int sZero;
   signed char *sd = (signed char *) smraw; unsigned char *sfl = (unsigned char *) smraw + g.l / 2;
   const uint32_t NCH = g.l / 2, L = g.l, tid = threadIdx.x, nth = blockDim.x;
   const mfft_move mv = moves[blockIdx.x / nbatch];
   const uint32_t bi = blockIdx.x % nbatch;
   const mfft_batch b = batch[bi];
   const cs_blk A = cs_block(slab, cw, g, mv.src_slot, b);
   limb_t *out = dst + ((uint64_t) dst_base[bi] + (uint64_t) mv.dst_pos * dst_stride) * g.pitch;
   const uint32_t per = (NCH + nth - 1) / nth;
   const uint32_t lo = (tid * per < NCH) ? tid * per : NCH, hi = (lo + per < NCH) ? lo + per : NCH;
   (void) nmoves;
   /* pass 1: absorb the previous chunk's carry word; the chunk goes to its place in dst */
   for (uint32_t i = lo; i < hi; i++)
   {
      limb_t x0, x1, r0, r1;
      ld2(x0, x1, A.x + 2 * i);
      const int32_t c = i ? A.c[i - 1] : 0;
      const int32_t k = add2(r0, r1, x0, x1, (limb_t)(int64_t) c, (limb_t)((int64_t) c >> 63));
      st2(out + 2 * i, r0, r1);
      sd[i] = (signed char)(k + (c >> 31));
      sfl[i] = (unsigned char)((((r0 & r1) == ~(limb_t) 0) ? 1u : 0u) | (((r0 | r1) == 0) ? 2u : 0u));
   }
   __syncthreads();
   /* pass 2: the run's transition function, block-wide inclusive scan (Hillis-Steele) */
   uint32_t F = 0x24u;                                   /* identity: -1 -> -1, 0 -> 0, +1 -> +1 (biased by 1) */
   for (uint32_t i = lo; i < hi; i++) F = tf_compose(cs_tf(sd[i], sfl[i]), F);
   sF[tid] = F;
   __syncthreads();
   for (uint32_t off = 1; off < nth; off <<= 1)
   {
      const uint32_t v = (tid >= off) ? tf_compose(sF[tid], sF[tid - off]) : sF[tid];
      __syncthreads();
      sF[tid] = v;
      __syncthreads();
   }
   uint32_t state = tid ? tf_apply(sF[tid - 1], 1u) : 1u;
   int64_t top = (int64_t) A.c[NCH - 1] + (int64_t) tf_apply(sF[nth - 1], 1u) - 1;
   /* pass 3: hand the carries on inside the run (almost always nothing to do) */
   for (uint32_t i = lo; i < hi; i++)
   {
      if (state != 1u)
      {
         limb_t x0, x1, r0, r1;
         ld2(x0, x1, out + 2 * i);
         const int64_t c = (int64_t) state - 1;
         (void) add2(r0, r1, x0, x1, (limb_t) c, (limb_t)(c >> 63));
         st2(out + 2 * i, r0, r1);
      }
      state = tf_apply(cs_tf(sd[i], sfl[i]), state);
   }
   __syncthreads();
   if (normalise)
   {  /* value = body + top B^l == body - top (mpn_normmod_2expp1, mul_fft.c:272-294) */
      for (int it = 0; it < 4; it++)
      {
         if (top == 0) break;
         if (top == 1)
         {
            if (tid == 0) sZero = 1;
            __syncthreads();
            bool z = true;
            for (uint32_t k = tid; k < L; k += nth) z = z && (out[k] == 0);
            if (!z) sZero = 0;
            __syncthreads();
            const int allz = sZero;
            __syncthreads();
            if (allz) break;
         }
         if (tid == 0)
         {
            int64_t c = -top, nt = 0;
            for (uint32_t pos = 0; c != 0; pos++)
            {
               if (pos == L) { nt = c; break; }
               const mfft_i128 acc = (mfft_i128) c + (mfft_i128)(mfft_u128) out[pos];
               out[pos] = (limb_t) acc; c = (int64_t)(acc >> 64);
            }
            sF[0] = (uint32_t)(int32_t) nt;
         }
         __syncthreads();
         top = (int64_t)(int32_t) sF[0];
         __syncthreads();
      }
   }
   if (tid == 0) out[L] = (limb_t) top;
}

#endif
