/* tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler of the SIMT emulator. */
#include "cuda_emu.h"

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace {
const size_t STACK = 256 * 1024;
struct fiber { ucontext_t ctx; char *stack; bool done; unsigned tid; };
std::vector<fiber> g_fibers;
ucontext_t g_sched;
fiber *g_cur = nullptr;
std::vector<emu_warp> g_warps;
emu_block_sync g_block;
emu_named_bar g_named[16];
void *g_smem = nullptr;
const std::function<void()> *g_body = nullptr;
unsigned long g_progress = 0;

void trampoline()
{
   (*g_body)();
   g_cur->done = true;
   /* a lane that leaves shrinks its warp and the block (whole warps leave together in our kernels) */
   emu_warp &w = g_warps[g_cur->tid >> 5];
   w.nlanes--; g_block.nthreads--;
   g_progress++;
   swapcontext(&g_cur->ctx, &g_sched);
}
}

void emu_yield() { swapcontext(&g_cur->ctx, &g_sched); }
void *emu_dyn_smem() { return g_smem; }
emu_warp *emu_cur_warp() { return &g_warps[g_cur->tid >> 5]; }
emu_block_sync *emu_cur_block() { return &g_block; }
emu_named_bar *emu_cur_named_bar(unsigned id) { return &g_named[id & 15]; }

void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body)
{
   if (g_fibers.size() < block)
   {
      size_t old = g_fibers.size();
      g_fibers.resize(block);
      for (size_t i = old; i < block; i++) g_fibers[i].stack = (char *) malloc(STACK);
   }
   std::vector<char> smem_buf(smem + 16);
   g_body = &body;
   gridDim = {grid, 1, 1}; blockDim = {block, 1, 1};
   for (unsigned b = 0; b < grid; b++)
   {
      const unsigned nwarps = (block + 31) / 32;
      g_warps.assign(nwarps, emu_warp());
      for (unsigned w = 0; w < nwarps; w++) { g_warps[w].gen = 0; g_warps[w].arrived = 0; g_warps[w].tag[0] = g_warps[w].tag[1] = 0xffffffffu; g_warps[w].nlanes = (w == nwarps - 1 && block % 32) ? block % 32 : 32; }
      g_block.gen = 0; g_block.arrived = 0; g_block.nthreads = block;
      memset(g_named, 0, sizeof g_named);
      g_smem = smem_buf.data();
      for (unsigned t = 0; t < block; t++)
      {
         fiber &f = g_fibers[t];
         f.done = false; f.tid = t;
         getcontext(&f.ctx);
         f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &g_sched;
         makecontext(&f.ctx, trampoline, 0);
      }
      unsigned live = block;
      while (live)
      {
         live = 0;
         const unsigned long before = g_progress;
         unsigned long yields = 0;
         for (unsigned t = 0; t < block; t++)
         {
            fiber &f = g_fibers[t];
            if (f.done) continue;
            g_cur = &f;
            threadIdx = {t, 0, 0}; blockIdx = {b, 0, 0};
            swapcontext(&g_sched, &f.ctx);
            if (!f.done) { live++; yields++; }
         }
         (void) before; (void) yields;
      }
   }
   g_body = nullptr;
}
