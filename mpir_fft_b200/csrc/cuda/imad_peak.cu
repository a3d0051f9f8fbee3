/* imad_peak.cu -- integer-multiply issue-rate microbenchmarks (the IMAD roofline denominator of
 * the pointwise kernel; MEASURED_PEAKS.json has only HBM and bf16).  Three instruction flavours,
 * each as 8 independent dependency chains per thread so that latency is hidden:
 *   0  IMAD.WIDE.U32          mad.wide.u32 -- NOTE: ptxas (sm_100a) never keeps the 64-bit addend in the
 *                             multiply: it emits IMAD.WIDE.U32 d, a, b, RZ plus IADD3 + IADD3.X on the
 *                             ALU pipe, so this measures multiplies with the accumulation on the other
 *                             pipe, not a fused multiply-add rate
 *   1  IMAD.WIDE.U32.X chain  carry in/out through predicates  (mad.lo.cc + madc.hi.cc, what the
 *                             schoolbook rows of k_pointwise compile to)
 *   2  IMAD (32-bit lo)       mad.lo.u32
 * Result: multiply-accumulates per second over the whole GPU. */
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_imad(uint32_t *out, uint32_t x, uint32_t y, int iters)
{
   uint64_t acc[8];
   uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
#pragma unroll
   for (int i = 0; i < 8; i++) acc[i] = i + threadIdx.x;
   uint32_t xs[8];
#pragma unroll
   for (int i = 0; i < 8; i++) xs[i] = x * (2 * i + 3) + threadIdx.x;
   for (int it = 0; it < iters; it++)
   {
      if (MODE == 0)
      {  /* sixteen DIFFERENT products per iteration (x[i]*b, x[i]*c), factors changing every
            iteration: nothing can be hoisted or shared between accumulators */
#pragma unroll
         for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(xs[i]), "r"(b));
#pragma unroll
         for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(xs[i]), "r"(a));
         a += 0x9e3779b9u; b ^= a;
      } else if (MODE == 1)
      {
         uint32_t *w = (uint32_t *) acc;
#pragma unroll
         for (int r = 0; r < 2; r++)
         {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(w[0]), "+r"(w[1]) : "r"(a), "r"(b));
#pragma unroll
            for (int i = 1; i < 8; i++)
               asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(w[2*i]), "+r"(w[2*i+1]) : "r"(a), "r"(b));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(a));
         }
      } else
      {
         uint32_t *w = (uint32_t *) acc;
#pragma unroll
         for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[i]) : "r"(a), "r"(b));
      }
   }
   uint64_t s = 0;
#pragma unroll
   for (int i = 0; i < 8; i++) s += acc[i];
   if (s == 0x1234567) out[0] = (uint32_t) s + a;
}

extern "C" double mfft_dev_imad_rate(int mode)
{
   int dev = 0, sms = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   uint32_t *d = NULL;
   if (cudaMalloc(&d, 64) != cudaSuccess) return -1.0;
   const int iters = 4096, grid = sms * 8, block = 256;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   float best = 1e30f;
   for (int rep = 0; rep < 5; rep++)
   {
      cudaEventRecord(e0);
      if (mode == 0) k_imad<0><<<grid, block>>>(d, 3, 5, iters);
      else if (mode == 1) k_imad<1><<<grid, block>>>(d, 3, 5, iters);
      else k_imad<2><<<grid, block>>>(d, 3, 5, iters);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return -1.0; }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
   const double ops = (double) grid * block * iters * 16.0;
   return ops / (best * 1e-3);
}
