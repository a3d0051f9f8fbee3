/* tile.h -- fused passes for the shared-memory tile executor (see tile.c) */
#ifndef MFFT_TILE_H
#define MFFT_TILE_H
#include "sched.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
   uint32_t     ntiles, max_npos, max_nops, nstages, npos_total, nops_total;
   uint32_t     nany;      /* ops the kernel has to decode at run time (MFFT_K_ANY): register-hungry path */
   uint32_t     nr4;       /* fused radix-4 units (hold 4 coefficients' chunks in registers) */
   mfft_tile   *tiles;     /* [ntiles] */
   uint32_t    *pos;       /* [npos_total] physical position | MFFT_TILE_LOAD | MFFT_TILE_STORE */
   mfft_tileop *ops;       /* [nops_total] */
   uint32_t    *stoff;     /* per tile nstages+1 offsets (local op index of each stage's first op) */
   uint32_t     nstoff;
} mfft_pass;

typedef struct { uint32_t npasses; mfft_pass *pass; } mfft_passes;

/* cut schedule s (position-based view) into passes whose tiles hold <= max_npos coefficients.
 * must_store (may be NULL): [S] flags of physical positions the LAST pass has to store even if no
 * op of its window writes them (the outputs of a transform whose last pass gathers into dst). */
/* live_out (may be NULL = all): [S] flags of the physical positions whose final value somebody needs
 * after the schedule; a position written by a pass is stored only if a later op reads it or it is live. */
int  mfft_passes_build(mfft_passes *P, const mfft_sched *s, uint32_t max_npos, const uint8_t *must_store, const uint8_t *live_out);
void mfft_passes_free(mfft_passes *P);
/* one pass from an explicit window of ops (see tile.c); last_read[p] = latest pstage that reads position p */
int  mfft_window_pass_build(mfft_pass *out, const mfft_op *ops, size_t nops, uint32_t S, uint32_t max_npos,
                            uint64_t NW, const uint32_t *last_read, const uint8_t *live_out, const uint8_t *must_store);

#ifdef __cplusplus
}
#endif
#endif
