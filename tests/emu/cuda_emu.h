/* tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal single-OS-thread SIMT emulator so that the *unmodified kernel source*
 * (mpir_fft_b200/csrc/cuda/mfft_kernels.cu) can be compiled with g++ and executed on a machine
 * without a GPU.  Every CUDA thread of a block is a ucontext fiber; warp collectives
 * (__ballot_sync, __shfl*_sync, __any/__all_sync, __syncwarp) and __syncthreads are rendezvous
 * points at which a fiber yields until all participants have arrived.  The CUDA runtime calls the
 * launch ABI uses are mapped onto malloc/memcpy.  The result, libmpirfft_emu.so, is loaded only
 * by the CPU test-suite to check kernel arithmetic and host logic; it is never shipped, and the
 * product library contains none of this (MFFT_EMU is undefined in the real build).
 */
#ifndef CUDA_EMU_H
#define CUDA_EMU_H
#ifndef MFFT_EMU
#error "cuda_emu.h is only for the CPU emulation build"
#endif

#include <ucontext.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __grid_constant__
#define __shared__ static

struct emu_dim3 { unsigned x, y, z; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

/* ---- runtime API subset ---- */
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 16); return *p ? cudaSuccess : 2; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **) p, n); }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{ for (size_t i = 0; i < h; i++) memmove((char *) d + i*dp, (const char *) s + i*sp, w); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef void *cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t) 1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t) 1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered, cudaMemoryTypeHost, cudaMemoryTypeDevice, cudaMemoryTypeManaged };
struct cudaPointerAttributes { cudaMemoryType type; };
/* the emulator has no pinned memory: everything counts as pageable unless EMU_ALL_PINNED is set (tests both paths) */
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *)
{ a->type = getenv("EMU_ALL_PINNED") ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

/* ---- fibers ---- */
void emu_yield();
void *emu_dyn_smem();
unsigned emu_lane_count_in_warp();      /* live lanes of the current warp */
void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body);

struct emu_warp {
   uint32_t gen; uint32_t arrived; uint32_t nlanes;
   uint64_t slot[2][32];
   uint32_t mask[2], tag[2];       /* which lanes took part in the collective of generation tag */
};
emu_warp *emu_cur_warp();
struct emu_block_sync { uint32_t gen, arrived, nthreads; };
emu_block_sync *emu_cur_block();

/* all live lanes of the warp deposit v, then everyone may read all deposits */
static inline const uint64_t *emu_warp_exchange(uint64_t v)
{
   emu_warp *w = emu_cur_warp();
   const uint32_t lane = threadIdx.x & 31, g = w->gen;
   if (w->tag[g & 1] != g) { w->tag[g & 1] = g; w->mask[g & 1] = 0; }
   w->mask[g & 1] |= 1u << lane;
   w->slot[g & 1][lane] = v;
   if (++w->arrived == w->nlanes) { w->arrived = 0; w->gen = g + 1; }
   else while (w->gen == g) emu_yield();
   return w->slot[g & 1];
}

static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_exchange(0); }
static inline unsigned emu_ballot(int pred, unsigned *members)
{
   emu_warp *w = emu_cur_warp();
   const uint32_t g = w->gen;                       /* generation this call will belong to */
   const uint64_t *s = emu_warp_exchange(pred ? 1 : 0);
   const unsigned m = w->mask[g & 1];
   unsigned r = 0;
   for (unsigned i = 0; i < 32; i++) if ((m >> i) & 1) r |= (unsigned)(s[i] & 1) << i;
   if (members) *members = m;
   return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) { return emu_ballot(pred, 0); }
static inline int __any_sync(unsigned, int pred) { return emu_ballot(pred, 0) != 0; }
static inline int __all_sync(unsigned, int pred) { unsigned m; unsigned r = emu_ballot(pred, &m); return r == m; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src)
{
   uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
   const uint64_t *s = emu_warp_exchange(raw);
   T out; uint64_t r = s[src & 31]; memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
   uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
   const uint64_t *s = emu_warp_exchange(raw);
   const unsigned lane = threadIdx.x & 31;
   T out; uint64_t r = (lane >= d) ? s[lane - d] : raw; memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
   uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
   const uint64_t *s = emu_warp_exchange(raw);
   const unsigned lane = threadIdx.x & 31;
   T out; uint64_t r = (lane + d < 32) ? s[lane + d] : raw; memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, unsigned m)
{
   uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
   const uint64_t *s = emu_warp_exchange(raw);
   const unsigned lane = threadIdx.x & 31;
   T out; uint64_t r = s[(lane ^ m) & 31]; memcpy(&out, &r, sizeof(T)); return out;
}
static inline void __threadfence_block() { }
static inline void __syncthreads()
{
   emu_block_sync *b = emu_cur_block();
   const uint32_t g = b->gen;
   if (++b->arrived == b->nthreads) { b->arrived = 0; b->gen = g + 1; }
   else while (b->gen == g) emu_yield();
}

/* named barrier (bar.sync id, nthreads): the nthreads threads that name the same id rendezvous */
struct emu_named_bar { uint32_t gen, arrived; };
emu_named_bar *emu_cur_named_bar(unsigned id);
static inline void emu_bar_sync(unsigned id, unsigned nthreads)
{
   emu_named_bar *b = emu_cur_named_bar(id);
   const uint32_t g = b->gen;
   if (++b->arrived == nthreads) { b->arrived = 0; b->gen = g + 1; }
   else while (b->gen == g) emu_yield();
}

/* mbarrier + bulk asynchronous copy (cp.async.bulk ... mbarrier::complete_tx): the emulated copy is
   done at issue; the barrier's phase completes when its arrival count is reached and the expected
   bytes have been delivered */
struct emu_mbar { uint32_t phase, count, arrived; int64_t pending; };
static inline void emu_mbar_try_complete(emu_mbar *m)
{ if (m->arrived >= m->count && m->pending == 0) { m->arrived = 0; m->phase ^= 1u; } }
static inline void emu_mbar_init(emu_mbar *m, uint32_t count) { m->phase = 0; m->count = count; m->arrived = 0; m->pending = 0; }
static inline void emu_mbar_arrive_expect_tx(emu_mbar *m, uint32_t bytes) { m->pending += bytes; m->arrived++; emu_mbar_try_complete(m); }
static inline void emu_bulk_copy(void *dst, const void *src, uint32_t bytes, emu_mbar *m)
{ memcpy(dst, src, bytes); m->pending -= bytes; emu_mbar_try_complete(m); }
static inline void emu_mbar_wait(emu_mbar *m, uint32_t parity) { while (m->phase == parity) emu_yield(); }

#define MFFT_LAUNCH(kern, grid, block, smem, stream, ...) \
   emu_launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
#define MFFT_LAUNCH_PDL(on, kern, grid, block, smem, stream, ...) \
   do { (void)(on); MFFT_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__); } while (0)
#define MFFT_DYN_SMEM(type, name) type *name = (type *) emu_dyn_smem()

#endif
