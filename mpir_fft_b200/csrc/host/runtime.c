/* runtime.c -- device binding, schedule upload/run, and the MFA transform driver (plain C99). */
#define _POSIX_C_SOURCE 200809L
#include "runtime.h"
#include "../../../include/mpirfft_b200.h"
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* One recursive lock serialises everything that touches library-global state: the device binding,
 * the plan caches of the drop-in symbols, plan creation / destruction and the issuing of a plan's
 * launches.  Recursive, because the drop-in symbols hold it across a whole call (cache lookup,
 * eviction, execution) and call the plan API, which takes it again. */
static pthread_mutex_t g_mu;
static pthread_once_t g_mu_once = PTHREAD_ONCE_INIT;
static int g_ready = 0;

static void mu_init(void)
{
   pthread_mutexattr_t a;
   pthread_mutexattr_init(&a);
   pthread_mutexattr_settype(&a, PTHREAD_MUTEX_RECURSIVE);
   pthread_mutex_init(&g_mu, &a);
   pthread_mutexattr_destroy(&a);
}

void mfft_lock(void) { pthread_once(&g_mu_once, mu_init); pthread_mutex_lock(&g_mu); }
void mfft_unlock(void) { pthread_mutex_unlock(&g_mu); }

void mfft_die(const char *fn, const char *fmt, ...)
{
   va_list ap;
   fprintf(stderr, "libmpirfft_b200: %s: ", fn);
   va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
   fprintf(stderr, "\n");
   fflush(stderr);
   abort();
}

int mfft_try_device(void)
{
   int rc = 0;
   mfft_lock();
   if (!g_ready)
   {
      const char *e = getenv("MPIRFFT_DEVICE");
      int dev = e ? atoi(e) : 0;
      if (mfft_dev_init(dev) != 0) rc = MPIRFFT_ENODEV; else g_ready = 1;
   }
   else if (mfft_dev_bind() != 0) rc = MPIRFFT_ENODEV;     /* the current device is per host thread */
   mfft_unlock();
   return rc;
}

void mfft_require_device(const char *fn)
{
   if (mfft_try_device() != 0)
      mfft_die(fn, "no usable CUDA device (%s); this library has no CPU fallback", mfft_dev_last_error());
}

int mpirfft_init(int device)
{
   int rc;
   mfft_lock();
   rc = mfft_dev_init(device);
   if (rc == 0) g_ready = 1;
   mfft_unlock();
   return rc == 0 ? MPIRFFT_OK : MPIRFFT_ENODEV;
}

int mpirfft_device_count(void) { return mfft_dev_count(); }
const char *mpirfft_last_error(void) { return mfft_dev_last_error(); }
const char *mpirfft_version(void) { return "mpirfft_b200 0.1 (sm_100a)"; }
double mfft_dev_imad_rate(int mode);
double mpirfft_measure_imad_rate(int mode) { return mfft_try_device() ? -1.0 : mfft_dev_imad_rate(mode); }

void mpirfft_set_pointwise_mode(int mode) { mfft_dev_pointwise_mode(mode); }

void mpirfft_profile_enable(int on) { mfft_dev_profile_enable(on); }
int  mpirfft_profile_read(double *ms, uint64_t *launches, double *bytes, int nclass)
{ return mfft_dev_profile_read(ms, launches, bytes, nclass) ? MPIRFFT_ENODEV : 0; }

uint64_t mpirfft_launch_count(void) { return mfft_dev_launch_count(); }
void mpirfft_launch_count_reset(void) { mfft_dev_launch_count_reset(); }

void *mpirfft_malloc_device(size_t bytes) { return mfft_try_device() ? NULL : mfft_dev_alloc(bytes); }
void  mpirfft_free_device(void *p) { mfft_dev_free(p); }
void *mpirfft_malloc_pinned(size_t bytes) { return mfft_try_device() ? NULL : mfft_host_alloc_pinned(bytes); }
void  mpirfft_free_pinned(void *p) { mfft_host_free_pinned(p); }
int   mpirfft_memcpy_h2d(void *d, const void *h, size_t bytes, void *stream) { return mfft_dev_h2d(d, h, bytes, stream) ? MPIRFFT_ENODEV : 0; }
int   mpirfft_memcpy_d2h(void *h, const void *d, size_t bytes, void *stream) { return mfft_dev_d2h(h, d, bytes, stream) ? MPIRFFT_ENODEV : 0; }
int   mpirfft_stream_sync(void *stream) { return mfft_dev_sync(stream) ? MPIRFFT_ENODEV : 0; }

void *mfft_upload(const void *h, size_t bytes)
{
   void *d = mfft_dev_alloc(bytes);
   if (!d) return NULL;
   if (bytes && (mfft_dev_h2d(d, h, bytes, NULL) != 0 || mfft_dev_sync(NULL) != 0)) { mfft_dev_free(d); return NULL; }
   return d;
}

int mfft_dsched_upload(mfft_dsched *ds, mfft_sched *s)
{
   ds->s = s; ds->d_ops = NULL;
   if (mfft_sched_finish(s) != 0) return MPIRFFT_ENOMEM;
   ds->d_ops = (mfft_op *) mfft_upload(s->ops, sizeof(mfft_op) * (s->nops ? s->nops : 1));
   return ds->d_ops ? 0 : MPIRFFT_ENODEV;
}

void mfft_dsched_free(mfft_dsched *ds)
{
   if (ds->d_ops) mfft_dev_free(ds->d_ops);
   if (ds->s) mfft_sched_free(ds->s);
   ds->d_ops = NULL; ds->s = NULL;
}

int mfft_dsched_run(const mfft_dsched *ds, limb_t *slab, const mfft_geom *g,
                    const mfft_batch *d_batch, uint32_t nbatch, void *stream)
{
   uint32_t st;
   for (st = 1; st <= ds->s->nstages; st++)
   {
      uint32_t lo = ds->s->stage_off[st - 1], hi = ds->s->stage_off[st], k;
      double blocks = 0;     /* blocks read + written by this stage: the algorithmic traffic */
      for (k = lo; k < hi; k++)
         blocks += 2 + (ds->s->ops[k].inB != MFFT_NONE) + (ds->s->ops[k].outT != MFFT_NONE);
      mfft_dev_profile_bytes(blocks * nbatch * 8.0 * (g->l + 1));
      if (hi > lo && mfft_dev_run_stage(slab, g, ds->d_ops + lo, hi - lo, d_batch, nbatch, stream) != 0)
         return MPIRFFT_ENODEV;
   }
   return 0;
}

/* consumed by the next forward plan (under the library lock) */
#define MFFT_ZERO_PROMISE_DEFAULT 0
static uint64_t g_promise_nz = 0;
void mfft_mfa_promise_zero_inputs(uint64_t nz)
{
   const char *e = getenv("MPIRFFT_ZERO_PROMISE");       /* 0 / 1; see MFFT_ZERO_PROMISE_DEFAULT */
   g_promise_nz = (e ? (e[0] != '0') : MFFT_ZERO_PROMISE_DEFAULT) ? nz : 0;
}

static uint32_t ilog2(uint64_t x) { uint32_t b = 0; while (((uint64_t)1 << b) < x) b++; return b; }

static void dpass_free(struct mfft_dpass *d, uint32_t n)
{
   uint32_t i;
   if (!d) return;
   for (i = 0; i < n; i++) { mfft_dev_free(d[i].d_tiles); mfft_dev_free(d[i].d_pos); mfft_dev_free(d[i].d_ops); mfft_dev_free(d[i].d_stoff); }
   free(d);
}

void mfft_mfa_free(mfft_mfa *m)
{
   if (m->col.s) m->h_col = NULL;     /* ownership moved to the uploaded schedule */
   if (m->row.s) m->h_row = NULL;
   mfft_dsched_free(&m->col); mfft_dsched_free(&m->row);
   if (m->h_col) mfft_sched_free(m->h_col);
   if (m->h_row) mfft_sched_free(m->h_row);
   if (m->col2.s) m->h_col2 = NULL;
   mfft_dsched_free(&m->col2);
   if (m->h_col2) mfft_sched_free(m->h_col2);
   dpass_free(m->dcol, m->pcol.npasses); dpass_free(m->drow, m->prow.npasses); dpass_free(m->dcol2, m->pcol2.npasses);
   mfft_passes_free(&m->pcol); mfft_passes_free(&m->prow); mfft_passes_free(&m->pcol2);
   mfft_dev_free(m->d_colb2); mfft_dev_free(m->d_dst_base2); free(m->h_colb2); free(m->h_dst_base2);
   mfft_dev_free(m->d_colb); mfft_dev_free(m->d_rowb); mfft_dev_free(m->d_moves); mfft_dev_free(m->d_dst_base);
   mfft_dev_free(m->d_dstpos);
   free(m->rows); free(m->h_colb); free(m->h_rowb); free(m->h_moves); free(m->h_dst_base);
   free(m->h_must_store); free(m->h_dstpos);
   memset(m, 0, sizeof(*m));
}

/* may the first column pass split while it loads?  Only if no later pass reads a block that
   nothing has written before (such a block would be an input the split kernel had to provide) */
static int passes_split_ok(const mfft_passes *P, uint32_t S)
{
   uint8_t *wr = (uint8_t *) calloc(S, 1); uint32_t pi, k; int ok = (wr != NULL);
   for (pi = 0; ok && pi < P->npasses; pi++)
   {
      const mfft_pass *p = &P->pass[pi];
      if (pi > 0)
         for (k = 0; k < p->npos_total; k++)
            if ((p->pos[k] & MFFT_TILE_LOAD) && !wr[p->pos[k] & MFFT_TILE_POSMASK]) ok = 0;
      /* with the split fused, the first pass loads from the operand and the slab holds only what a
         pass has STORED */
      for (k = 0; k < p->npos_total; k++)
         if (p->pos[k] & MFFT_TILE_STORE) wr[p->pos[k] & MFFT_TILE_POSMASK] = 1;
   }
   free(wr);
   return ok;
}

int mfft_mfa_plan(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                  uint32_t final_shift, int normalise, int mode)
{
   uint64_t NW = n*w, n2, i, j;
   mfft_sched *cs, *rs, *last; mfft_batch *colb, *rowb; mfft_move *mv; uint32_t *dstb;
   int rc = MPIRFFT_EINVAL;
   const char *env = getenv("MPIRFFT_UNFUSED");

   memset(m, 0, sizeof(*m));
   if (n == 0 || (n & (n - 1)) || w == 0 || (NW % 64) || n1 < 2 || (n1 & (n1 - 1)) || n1 > n) return MPIRFFT_EINVAL;
   n2 = 2*n/n1;
   if (n2 < 2) return MPIRFFT_EINVAL;
   m->inverse = inverse; m->truncated = (trunc != 0);
   m->n = n; m->w = w; m->n1 = n1; m->n2 = n2; m->N = 2*n;
   m->l = (uint32_t)(NW/64); m->pitch = mfft_pitch(m->l);
   m->depth1 = ilog2(n2); m->depth2 = ilog2(n1);
   m->final_shift = (uint32_t)(final_shift % (2*NW)); m->normalise = normalise;
   m->fused = (mode == 0) && mfft_dev_tiles_supported(m->l) && !(env && env[0] == '1');
   if (trunc)
   {  /* trunc must be a multiple of 2*n1 (mul_fft.c:2209-2211, 3200) */
      if (trunc % (2*n1) || trunc > 2*n) return MPIRFFT_EINVAL;
      m->trunc_rows = trunc/n1;
   } else m->trunc_rows = n2;

   /* the valid rows: revbin(s), s < trunc_rows (2392-2394, 2942-2944) */
   m->nrows = (uint32_t) m->trunc_rows;
   m->rows = (uint32_t *) malloc(sizeof(uint32_t) * m->nrows);
   m->h_col = cs = mfft_sched_new((uint32_t) n2, NW);
   m->h_row = rs = mfft_sched_new((uint32_t) n1, NW);
   m->h_colb = colb = (mfft_batch *) calloc(n1, sizeof(mfft_batch));
   m->h_rowb = rowb = (mfft_batch *) calloc(m->nrows, sizeof(mfft_batch));
   if (!m->rows || !cs || !rs || !colb || !rowb) { rc = MPIRFFT_ENOMEM; goto fail; }
   for (i = 0; i < m->nrows; i++) m->rows[i] = (uint32_t) mfft_revbin(i, m->depth1);

   m->gcol.S = (uint32_t) n2; m->gcol.slot_stride = (uint32_t) n1; m->gcol.half_blocks = m->N;
   m->gcol.l = m->l; m->gcol.pitch = m->pitch;
   m->grow.S = (uint32_t) n1; m->grow.slot_stride = 1; m->grow.half_blocks = m->N;
   m->grow.l = m->l; m->grow.pitch = m->pitch;
   m->ncolb = (uint32_t) n1; m->nrowb = m->nrows;

   if (!inverse)
   {
      /* rows that the split leaves zero (position-view plans only: the ping-pong view keeps every op) */
      if (g_promise_nz && m->fused && (g_promise_nz + n1 - 1)/n1 < n2 && mfft_sched_zero_from(cs, (uint32_t)((g_promise_nz + n1 - 1)/n1)) != 0)
      { rc = MPIRFFT_ENOMEM; goto fail; }
      /* column FFTs with the fused twist (2374-2379), then the row relabel (2380-2389) */
      if (mfft_sched_emit(cs, trunc ? MFFT_T_FFT_TRUNC : MFFT_T_FFT, 0, 1, n2/2, w*n1, w, 0, 1, m->trunc_rows) != 0) goto fail;
      mfft_sched_revbin(cs, 0, 1, m->depth1);
      for (j = 0; j < n1; j++) { colb[j].base = (uint32_t) j; colb[j].parity = 0; colb[j].col = (uint32_t) j; }
      /* row FFTs on the valid rows (2392-2395), then the in-row relabel (2397-2405) */
      if (mfft_sched_emit(rs, MFFT_T_FFT, 0, 1, n1/2, w*n2, 0, 0, 0, 0) != 0) goto fail;
      mfft_sched_revbin(rs, 0, 1, m->depth2);
      for (i = 0; i < m->nrows; i++)
      {  /* where the column pass left logical row rows[i]: slot (ping-pong) or physical row (fused) */
         uint32_t slot = m->fused ? cs->phys[m->rows[i]] : cs->slot[m->rows[i]];
         rowb[i].base = (uint32_t)((slot % n2) * n1); rowb[i].parity = (uint32_t)(slot / n2); rowb[i].col = 0;
      }
      /* finalize: logical (row i, col j) -> dst block i*n1 + j */
      m->nmoves = (uint32_t) n1; m->ndst = m->nrows;
      m->h_moves = mv = (mfft_move *) calloc(m->nmoves, sizeof(mfft_move));
      m->h_dst_base = dstb = (uint32_t *) calloc(m->ndst, sizeof(uint32_t));
      if (!mv || !dstb) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (j = 0; j < n1; j++) { mv[j].src_slot = rs->slot[j]; mv[j].dst_pos = (uint32_t) j; }
      for (i = 0; i < m->nrows; i++) dstb[i] = (uint32_t)(m->rows[i] * n1);
      m->dst_stride = 1;
      last = rs;
   } else
   {
      /* in-row relabel then row IFFTs on the valid rows (2942-2956) */
      mfft_sched_revbin(rs, 0, 1, m->depth2);
      if (mfft_sched_emit(rs, MFFT_T_IFFT, 0, 1, n1/2, w*n2, 0, 0, 0, 0) != 0) goto fail;
      for (i = 0; i < m->nrows; i++) { rowb[i].base = (uint32_t)(m->rows[i] * n1); rowb[i].parity = 0; rowb[i].col = 0; }
      /* row relabel then column IFFTs with the twist removed (2959-2974) */
      mfft_sched_revbin(cs, 0, 1, m->depth1);
      if (mfft_sched_emit(cs, trunc ? MFFT_T_IFFT_TRUNC : MFFT_T_IFFT, 0, 1, n2/2, w*n1, w, 0, 1, m->trunc_rows) != 0) goto fail;
      for (j = 0; j < n1; j++)
      {
         uint32_t slot = m->fused ? rs->phys[j] : rs->slot[j];
         colb[j].base = (uint32_t)(slot % n1); colb[j].parity = (uint32_t)(slot / n1); colb[j].col = (uint32_t) j;
      }
      /* finalize: logical (row r < trunc_rows, col j) -> dst block r*n1 + j */
      m->nmoves = (uint32_t) m->trunc_rows; m->ndst = (uint32_t) n1;
      m->h_moves = mv = (mfft_move *) calloc(m->nmoves, sizeof(mfft_move));
      m->h_dst_base = dstb = (uint32_t *) calloc(m->ndst, sizeof(uint32_t));
      if (!mv || !dstb) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (i = 0; i < m->nmoves; i++) { mv[i].src_slot = cs->slot[i]; mv[i].dst_pos = (uint32_t) i; }
      for (j = 0; j < n1; j++) dstb[j] = (uint32_t) j;
      m->dst_stride = (uint32_t) n1;
      last = cs;
   }
   if (m->fused)
   {  /* the last schedule's final pass gathers the logical outputs into dst; a final scaling is
         one more in-place op per output */
      uint32_t nout = m->nmoves, S = last->S, k;
      uint32_t pmax = mfft_dev_tiles_max_npos(m->l);
      m->h_must_store = (uint8_t *) calloc(S, 1);
      m->h_dstpos = (uint32_t *) malloc(sizeof(uint32_t) * S);
      if (!m->h_must_store || !m->h_dstpos) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (k = 0; k < S; k++) m->h_dstpos[k] = MFFT_NONE;
      if (m->final_shift)
         for (k = 0; k < nout; k++)
            mfft_sched_emit_op(last, k, MFFT_NONE, k, 1, m->final_shift, 0, 0, MFFT_NONE, 0, 0, 0, 0);
      for (k = 0; k < nout; k++) { m->h_must_store[last->phys[k]] = 1; m->h_dstpos[last->phys[k]] = k; }
      {  /* what the first schedule has to leave behind: forward, the rows the row pass will read
            (2392-2394); inverse, every column of the valid rows (all positions of the row schedule) */
         uint8_t *live_first = NULL;
         if (!inverse)
         {
            live_first = (uint8_t *) calloc(cs->S, 1);
            if (!live_first) { rc = MPIRFFT_ENOMEM; goto fail; }
            for (i = 0; i < m->nrows; i++) live_first[cs->phys[m->rows[i]]] = 1;
         }
         if (mfft_passes_build(&m->pcol, cs, pmax, last == cs ? m->h_must_store : NULL, last == cs ? m->h_must_store : live_first) != 0 ||
             mfft_passes_build(&m->prow, rs, pmax, last == rs ? m->h_must_store : NULL, last == rs ? m->h_must_store : NULL) != 0)
         { free(live_first); rc = MPIRFFT_ENOMEM; goto fail; }
         free(live_first);
      }
   }
   if (m->fused && !inverse) m->fuse_split_ok = passes_split_ok(&m->pcol, cs->S);
   if (mfft_sched_finish(cs) != 0 || mfft_sched_finish(rs) != 0) { rc = MPIRFFT_ENOMEM; goto fail; }
   return 0;
fail:
   mfft_mfa_free(m);
   return rc;
}

static struct mfft_dpass *dpass_upload(const mfft_passes *P)
{
   uint32_t i;
   struct mfft_dpass *d = (struct mfft_dpass *) calloc(P->npasses ? P->npasses : 1, sizeof(*d));
   if (!d) return NULL;
   for (i = 0; i < P->npasses; i++)
   {
      const mfft_pass *p = &P->pass[i];
      d[i].d_tiles = (mfft_tile *) mfft_upload(p->tiles, sizeof(mfft_tile) * (p->ntiles ? p->ntiles : 1));
      d[i].d_pos = (uint32_t *) mfft_upload(p->pos, sizeof(uint32_t) * (p->npos_total ? p->npos_total : 1));
      d[i].d_ops = (mfft_tileop *) mfft_upload(p->ops, sizeof(mfft_tileop) * (p->nops_total ? p->nops_total : 1));
      d[i].d_stoff = (uint32_t *) mfft_upload(p->stoff, sizeof(uint32_t) * (p->nstoff ? p->nstoff : 1));
      if (!d[i].d_tiles || !d[i].d_pos || !d[i].d_ops || !d[i].d_stoff) { dpass_free(d, P->npasses); return NULL; }
   }
   return d;
}

int mfft_mfa_upload(mfft_mfa *m)
{
   if (mfft_dsched_upload(&m->col, m->h_col) != 0) return MPIRFFT_ENODEV;
   if (mfft_dsched_upload(&m->row, m->h_row) != 0) return MPIRFFT_ENODEV;
   m->d_colb = (mfft_batch *) mfft_upload(m->h_colb, sizeof(mfft_batch) * m->ncolb);
   m->d_rowb = (mfft_batch *) mfft_upload(m->h_rowb, sizeof(mfft_batch) * m->nrowb);
   m->d_moves = (mfft_move *) mfft_upload(m->h_moves, sizeof(mfft_move) * m->nmoves);
   m->d_dst_base = (uint32_t *) mfft_upload(m->h_dst_base, sizeof(uint32_t) * m->ndst);
   if (!m->d_colb || !m->d_rowb || !m->d_moves || !m->d_dst_base) return MPIRFFT_ENODEV;
   if (m->nclass == 2)
   {
      if (mfft_dsched_upload(&m->col2, m->h_col2) != 0) return MPIRFFT_ENODEV;
      m->d_colb2 = (mfft_batch *) mfft_upload(m->h_colb2, sizeof(mfft_batch) * m->ncolb2);
      if (!m->d_colb2) return MPIRFFT_ENODEV;
      if (m->h_dst_base2)
      {
         m->d_dst_base2 = (uint32_t *) mfft_upload(m->h_dst_base2, sizeof(uint32_t) * m->ndst2);
         if (!m->d_dst_base2) return MPIRFFT_ENODEV;
      }
   }
   if (m->fused)
   {
      uint32_t S = m->inverse ? m->gcol.S : m->grow.S;
      m->dcol = dpass_upload(&m->pcol); m->drow = dpass_upload(&m->prow);
      m->d_dstpos = (uint32_t *) mfft_upload(m->h_dstpos, sizeof(uint32_t) * S);
      if (!m->dcol || !m->drow || !m->d_dstpos) return MPIRFFT_ENODEV;
      if (m->nclass == 2 && !(m->dcol2 = dpass_upload(&m->pcol2))) return MPIRFFT_ENODEV;
   }
   return 0;
}

int mfft_mfa_build(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                   uint32_t final_shift, int normalise)
{
   int rc = mfft_mfa_plan(m, inverse, n, w, n1, trunc, final_shift, normalise, 0);
   g_promise_nz = 0;
   if (rc != 0) return rc;
   if ((rc = mfft_mfa_upload(m)) != 0) { mfft_mfa_free(m); return rc; }
   return 0;
}


/* ---- sqrt2 MFA (FFT/IFFT_radix2_mfa_truncate_sqrt2, mul_fft.c:2212-2355, 2593-2750) ------------------ */
int mfft_mfa_plan_sqrt2(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                        uint32_t final_shift, int normalise, int mode)
{
   uint64_t NW = n*w, n2, trunc2, i, j; uint32_t cl, k;
   mfft_sched *cs[2] = { NULL, NULL }, *rs, *last; mfft_batch *rowb; mfft_move *mv;
   int rc = MPIRFFT_EINVAL;
   const char *env = getenv("MPIRFFT_UNFUSED");

   memset(m, 0, sizeof(*m));
   if (n == 0 || (n & (n - 1)) || w == 0 || (NW % 64) || (NW % 4) || n1 < 2 || (n1 & (n1 - 1)) || n1 > n) return MPIRFFT_EINVAL;
   n2 = 2*n/n1;
   if (n2 < 4) return MPIRFFT_EINVAL;
   /* trunc: a multiple of 2*n1 in (2n, 4n]  (mul_fft.c:2209-2211) */
   if (trunc % (2*n1) || trunc <= 2*n || trunc > 4*n) return MPIRFFT_EINVAL;
   trunc2 = (trunc - 2*n)/n1;
   m->sqrt2 = 1; m->nclass = (w & 1) ? 2 : 1; m->trunc2 = trunc2;
   m->inverse = inverse; m->truncated = 1;
   m->n = n; m->w = w; m->n1 = n1; m->n2 = n2; m->N = 4*n;
   m->l = (uint32_t)(NW/64); m->pitch = mfft_pitch(m->l);
   m->depth1 = ilog2(n2); m->depth2 = ilog2(n1);
   m->final_shift = (uint32_t)(final_shift % (2*NW)); m->normalise = normalise;
   m->fused = (mode == 0) && mfft_dev_tiles_supported(m->l) && !(env && env[0] == '1');
   m->trunc_rows = n2 + trunc2;

   /* the valid rows, as positions of a column: all first-half rows, then revbin(s) of the second half */
   m->nrows = (uint32_t)(n2 + trunc2);
   m->rows = (uint32_t *) malloc(sizeof(uint32_t) * m->nrows);
   m->h_row = rs = mfft_sched_new((uint32_t) n1, NW);
   m->h_rowb = rowb = (mfft_batch *) calloc(m->nrows, sizeof(mfft_batch));
   if (!m->rows || !rs || !rowb) { rc = MPIRFFT_ENOMEM; goto fail; }
   for (i = 0; i < n2; i++) m->rows[i] = (uint32_t) i;
   for (i = 0; i < trunc2; i++) m->rows[n2 + i] = (uint32_t)(n2 + mfft_revbin(i, m->depth1));

   m->gcol.S = (uint32_t)(2*n2); m->gcol.slot_stride = (uint32_t) n1; m->gcol.half_blocks = m->N;
   m->gcol.l = m->l; m->gcol.pitch = m->pitch;
   m->grow.S = (uint32_t) n1; m->grow.slot_stride = 1; m->grow.half_blocks = m->N;
   m->grow.l = m->l; m->grow.pitch = m->pitch;
   m->nrowb = m->nrows;

   for (cl = 0; cl < (uint32_t) m->nclass; cl++)
   {
      const uint32_t ncb = (uint32_t)(m->nclass == 2 ? n1/2 : n1);
      mfft_batch *cb = (mfft_batch *) calloc(ncb, sizeof(mfft_batch));
      cs[cl] = mfft_sched_new((uint32_t)(2*n2), NW);
      if (cl == 0) { m->h_col = cs[0]; m->h_colb = cb; m->ncolb = ncb; } else { m->h_col2 = cs[1]; m->h_colb2 = cb; m->ncolb2 = ncb; }
      if (!cs[cl] || !cb) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (j = 0; j < ncb; j++)
      {  /* column = 2 col + par (two classes) or col */
         cb[j].base = (uint32_t)(m->nclass == 2 ? 2*j + cl : j); cb[j].parity = 0; cb[j].col = (uint32_t) j;
      }
   }
   if (!inverse)
   {
      for (cl = 0; cl < (uint32_t) m->nclass; cl++)
         if (g_promise_nz && m->fused && (g_promise_nz + n1 - 1)/n1 < 2*n2 && mfft_sched_zero_from(cs[cl], (uint32_t)((g_promise_nz + n1 - 1)/n1)) != 0)
         { rc = MPIRFFT_ENOMEM; goto fail; }
      for (cl = 0; cl < (uint32_t) m->nclass; cl++)
         if (mfft_sched_emit_sqrt2_cols(cs[cl], 0, n2, n1, w, trunc2, m->nclass == 2 ? (int) cl : -1, !m->fused) != 0) goto fail;
      /* row FFTs on the valid rows, then the in-row relabel (2289-2303, 2341-2354) */
      if (mfft_sched_emit(rs, MFFT_T_FFT, 0, 1, n1/2, w*n2, 0, 0, 0, 0) != 0) goto fail;
      mfft_sched_revbin(rs, 0, 1, m->depth2);
      for (i = 0; i < m->nrows; i++)
      {
         uint32_t slot = m->fused ? cs[0]->phys[m->rows[i]] : cs[0]->slot[m->rows[i]];
         if (m->nclass == 2 && slot != (m->fused ? cs[1]->phys[m->rows[i]] : cs[1]->slot[m->rows[i]])) { rc = MPIRFFT_EINVAL; goto fail; }
         rowb[i].base = (uint32_t)((slot % (2*n2)) * n1); rowb[i].parity = (uint32_t)(slot / (2*n2)); rowb[i].col = 0;
      }
      m->nmoves = (uint32_t) n1; m->ndst = m->nrows;
      m->h_moves = mv = (mfft_move *) calloc(m->nmoves, sizeof(mfft_move));
      m->h_dst_base = (uint32_t *) calloc(m->ndst, sizeof(uint32_t));
      if (!mv || !m->h_dst_base) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (j = 0; j < n1; j++) { mv[j].src_slot = rs->slot[j]; mv[j].dst_pos = (uint32_t) j; }
      for (i = 0; i < m->nrows; i++) m->h_dst_base[i] = (uint32_t)(m->rows[i] * n1);
      m->dst_stride = 1;
      last = rs;
   } else
   {
      mfft_sched_revbin(rs, 0, 1, m->depth2);
      if (mfft_sched_emit(rs, MFFT_T_IFFT, 0, 1, n1/2, w*n2, 0, 0, 0, 0) != 0) goto fail;
      for (i = 0; i < m->nrows; i++) { rowb[i].base = (uint32_t)(m->rows[i] * n1); rowb[i].parity = 0; rowb[i].col = 0; }
      for (cl = 0; cl < (uint32_t) m->nclass; cl++)
      {
         mfft_batch *cb = cl ? m->h_colb2 : m->h_colb; uint32_t ncb = cl ? m->ncolb2 : m->ncolb;
         if (mfft_sched_emit_sqrt2_cols(cs[cl], 1, n2, n1, w, trunc2, m->nclass == 2 ? (int) cl : -1, !m->fused) != 0) goto fail;
         for (j = 0; j < ncb; j++)
         {  /* where the row pass left this column */
            uint32_t c = cb[j].base, slot = m->fused ? rs->phys[c] : rs->slot[c];
            cb[j].base = (uint32_t)(slot % n1); cb[j].parity = (uint32_t)(slot / n1);
         }
      }
      /* outputs: logical positions 0 .. n2+trunc2-1 of every column -> dst block pos*n1 + column */
      m->nmoves = m->nrows;
      m->h_moves = mv = (mfft_move *) calloc(m->nmoves, sizeof(mfft_move));
      if (!mv) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (i = 0; i < m->nmoves; i++)
      {
         mv[i].src_slot = cs[0]->slot[i]; mv[i].dst_pos = (uint32_t) i;
         /* both column classes must leave their outputs in the same places (the view that is executed) */
         if (m->nclass == 2 && (m->fused ? cs[1]->phys[i] != cs[0]->phys[i] : cs[1]->slot[i] != cs[0]->slot[i])) { rc = MPIRFFT_EINVAL; goto fail; }
      }
      m->ndst = m->ncolb; m->h_dst_base = (uint32_t *) calloc(m->ndst, sizeof(uint32_t));
      if (!m->h_dst_base) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (j = 0; j < m->ncolb; j++) m->h_dst_base[j] = (uint32_t)(m->nclass == 2 ? 2*j : j);
      if (m->nclass == 2)
      {
         m->ndst2 = m->ncolb2; m->h_dst_base2 = (uint32_t *) calloc(m->ndst2, sizeof(uint32_t));
         if (!m->h_dst_base2) { rc = MPIRFFT_ENOMEM; goto fail; }
         for (j = 0; j < m->ncolb2; j++) m->h_dst_base2[j] = (uint32_t)(2*j + 1);
      }
      m->dst_stride = (uint32_t) n1;
      last = cs[0];
   }
   if (m->fused)
   {
      uint32_t nout = m->nmoves, S = last->S;
      uint32_t pmax = mfft_dev_tiles_max_npos(m->l);
      uint8_t *live_first = NULL;
      m->h_must_store = (uint8_t *) calloc(S, 1);
      m->h_dstpos = (uint32_t *) malloc(sizeof(uint32_t) * S);
      if (!m->h_must_store || !m->h_dstpos) { rc = MPIRFFT_ENOMEM; goto fail; }
      for (k = 0; k < S; k++) m->h_dstpos[k] = MFFT_NONE;
      if (m->final_shift)
         for (cl = 0; cl < (uint32_t)(last == rs ? 1 : m->nclass); cl++)
            for (k = 0; k < nout; k++)
               mfft_sched_emit_op(last == rs ? rs : cs[cl], k, MFFT_NONE, k, 1, m->final_shift, 0, 0, MFFT_NONE, 0, 0, 0, 0);
      for (k = 0; k < nout; k++) { m->h_must_store[last->phys[k]] = 1; m->h_dstpos[last->phys[k]] = k; }
      if (!inverse)
      {
         live_first = (uint8_t *) calloc(cs[0]->S, 1);
         if (!live_first) { rc = MPIRFFT_ENOMEM; goto fail; }
         for (i = 0; i < m->nrows; i++) live_first[cs[0]->phys[m->rows[i]]] = 1;
      }
      rc = MPIRFFT_ENOMEM;
      if (mfft_passes_build(&m->pcol, cs[0], pmax, last == rs ? NULL : m->h_must_store, last == rs ? live_first : m->h_must_store) != 0 ||
          (m->nclass == 2 && mfft_passes_build(&m->pcol2, cs[1], pmax, last == rs ? NULL : m->h_must_store, last == rs ? live_first : m->h_must_store) != 0) ||
          mfft_passes_build(&m->prow, rs, pmax, last == rs ? m->h_must_store : NULL, last == rs ? m->h_must_store : NULL) != 0)
      { free(live_first); goto fail; }
      free(live_first);
      if (!inverse)
         m->fuse_split_ok = passes_split_ok(&m->pcol, cs[0]->S) && (m->nclass == 1 || passes_split_ok(&m->pcol2, cs[1]->S));
   }
   if (mfft_sched_finish(cs[0]) != 0 || (m->nclass == 2 && mfft_sched_finish(cs[1]) != 0) || mfft_sched_finish(rs) != 0) { rc = MPIRFFT_ENOMEM; goto fail; }
   return 0;
fail:
   mfft_mfa_free(m);
   return rc;
}

int mfft_mfa_build_sqrt2(mfft_mfa *m, int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc,
                         uint32_t final_shift, int normalise)
{
   int rc = mfft_mfa_plan_sqrt2(m, inverse, n, w, n1, trunc, final_shift, normalise, 0);
   g_promise_nz = 0;
   if (rc != 0) return rc;
   if ((rc = mfft_mfa_upload(m)) != 0) { mfft_mfa_free(m); return rc; }
   return 0;
}

/* accessors for the CPU-side schedule tests (tests/schedsim.py) */
mfft_mfa *mfft_mfa_debug_new(int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc)
{
   mfft_mfa *m = (mfft_mfa *) calloc(1, sizeof(*m));
   if (m && mfft_mfa_plan(m, inverse, n, w, n1, trunc, 0, 0, 1) != 0) { free(m); return NULL; }
   return m;
}
void mfft_mfa_debug_free(mfft_mfa *m) { if (m) { mfft_mfa_free(m); free(m); } }
mfft_mfa *mfft_mfa_debug_new_sqrt2(int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc)
{
   mfft_mfa *m = (mfft_mfa *) calloc(1, sizeof(*m));
   if (m && mfft_mfa_plan_sqrt2(m, inverse, n, w, n1, trunc, 0, 0, 1) != 0) { free(m); return NULL; }
   return m;
}
/* which: 0 column schedule (class 0), 1 row schedule, 2 column schedule of the odd columns (sqrt2, odd w) */
mfft_sched *mfft_mfa_debug_sched(mfft_mfa *m, int which) { return which == 2 ? m->h_col2 : which ? m->h_row : m->h_col; }
mfft_batch *mfft_mfa_debug_batch(mfft_mfa *m, int which, uint32_t *count)
{ *count = which == 2 ? m->ncolb2 : which ? m->nrowb : m->ncolb; return which == 2 ? m->h_colb2 : which ? m->h_rowb : m->h_colb; }
uint32_t *mfft_mfa_debug_dst_base2(mfft_mfa *m, uint32_t *count) { *count = m->ndst2; return m->h_dst_base2; }
mfft_move *mfft_mfa_debug_moves(mfft_mfa *m, uint32_t *count, uint32_t *dst_stride)
{ *count = m->nmoves; *dst_stride = m->dst_stride; return m->h_moves; }
uint32_t *mfft_mfa_debug_dst_base(mfft_mfa *m, uint32_t *count) { *count = m->ndst; return m->h_dst_base; }
uint32_t *mfft_mfa_debug_rows(mfft_mfa *m, uint32_t *count) { *count = m->nrows; return m->rows; }

/* algorithmic traffic of one MFA pass by SURVEY 8(d): one read and one write of the live slab,
 * 2 T S bytes with T = trunc live coefficients and S = 8 (l+1) -- NOT what the executor happens to
 * move (rows the truncation synthesises, operand bytes of a split-fused first pass), so that
 * bytes / time is comparable across implementations */
static double pass_bytes(const mfft_mfa *m)
{
   return 2.0 * (double)(m->sqrt2 ? m->nrows : m->trunc_rows) * (double) m->n1 * 8.0 * (m->l + 1);
}

static int run_passes(const mfft_mfa *m, const mfft_passes *P, const struct mfft_dpass *d, limb_t *slab,
                      const mfft_geom *g, const mfft_batch *d_batch, const mfft_batch *h_batch, uint32_t nbatch, limb_t *dst,
                      const mfft_split *split, void *stream, const uint32_t *d_dst_base)
{
   uint32_t i;
   for (i = 0; i < P->npasses; i++)
   {
      const mfft_pass *p = &P->pass[i];
      const int lastp = (i + 1 == P->npasses) && dst != NULL;
      mfft_dev_profile_bytes(pass_bytes(m));
      if (mfft_dev_run_tiles(slab, g, d[i].d_tiles, p->ntiles, d[i].d_pos, d[i].d_ops, p->max_npos, p->max_nops, d_batch, nbatch,
                             lastp ? dst : NULL, m->d_dstpos, d_dst_base, m->dst_stride,
                             lastp ? m->normalise : 0, d[i].d_stoff, (5*p->nany > p->nops_total) || p->nr4, p->tiles, p->pos, p->stoff, h_batch, (i == 0) ? split : NULL, stream) != 0) return MPIRFFT_ENODEV;
   }
   return 0;
}

/* forward transform straight from the operand: the first column pass splits while it loads
 * (fused plans whose first pass reads every live input block exactly once); returns 1 if the
 * caller still has to run the split kernel */
int mfft_mfa_can_fuse_split(const mfft_mfa *m) { return m->fused && !m->inverse && m->pcol.npasses > 0 && m->fuse_split_ok; }

static int mfa_exec(const mfft_mfa *m, limb_t *slab, limb_t *dst, const mfft_split *split, void *stream);

int mfft_mfa_exec_split(const mfft_mfa *m, limb_t *slab, limb_t *dst, const limb_t *src, uint64_t nlimbs,
                        uint64_t bits, uint64_t ncoef, void *stream)
{
   mfft_split sp;
   if (!mfft_mfa_can_fuse_split(m)) return MPIRFFT_EINVAL;
   sp.src = src; sp.nlimbs = nlimbs; sp.bits = bits; sp.ncoef = ncoef;
   return mfa_exec(m, slab, dst, &sp, stream);
}

int mfft_mfa_exec(const mfft_mfa *m, limb_t *slab, limb_t *dst, void *stream) { return mfa_exec(m, slab, dst, NULL, stream); }

static int mfa_exec(const mfft_mfa *m, limb_t *slab, limb_t *dst, const mfft_split *split, void *stream)
{
   int rc;
   if (m->fused)
   {
      if (!m->inverse)
      {
         if ((rc = run_passes(m, &m->pcol, m->dcol, slab, &m->gcol, m->d_colb, m->h_colb, m->ncolb, NULL, split, stream, NULL)) != 0) return rc;
         if (m->nclass == 2 && (rc = run_passes(m, &m->pcol2, m->dcol2, slab, &m->gcol, m->d_colb2, m->h_colb2, m->ncolb2, NULL, split, stream, NULL)) != 0) return rc;
         return run_passes(m, &m->prow, m->drow, slab, &m->grow, m->d_rowb, m->h_rowb, m->nrowb, dst, NULL, stream, m->d_dst_base);
      }
      if ((rc = run_passes(m, &m->prow, m->drow, slab, &m->grow, m->d_rowb, m->h_rowb, m->nrowb, NULL, NULL, stream, NULL)) != 0) return rc;
      if ((rc = run_passes(m, &m->pcol, m->dcol, slab, &m->gcol, m->d_colb, m->h_colb, m->ncolb, dst, NULL, stream, m->d_dst_base)) != 0) return rc;
      if (m->nclass == 2) return run_passes(m, &m->pcol2, m->dcol2, slab, &m->gcol, m->d_colb2, m->h_colb2, m->ncolb2, dst, NULL, stream, m->d_dst_base2);
      return 0;
   }
   if (!m->inverse)
   {
      if ((rc = mfft_dsched_run(&m->col, slab, &m->gcol, m->d_colb, m->ncolb, stream)) != 0) return rc;
      if (m->nclass == 2 && (rc = mfft_dsched_run(&m->col2, slab, &m->gcol, m->d_colb2, m->ncolb2, stream)) != 0) return rc;
      if ((rc = mfft_dsched_run(&m->row, slab, &m->grow, m->d_rowb, m->nrowb, stream)) != 0) return rc;
      if (mfft_dev_finalize(dst, m->dst_stride, m->d_dst_base, slab, &m->grow, m->d_moves, m->nmoves,
                            m->d_rowb, m->nrowb, m->final_shift, m->normalise, stream) != 0) return MPIRFFT_ENODEV;
   } else
   {
      if ((rc = mfft_dsched_run(&m->row, slab, &m->grow, m->d_rowb, m->nrowb, stream)) != 0) return rc;
      if ((rc = mfft_dsched_run(&m->col, slab, &m->gcol, m->d_colb, m->ncolb, stream)) != 0) return rc;
      if (mfft_dev_finalize(dst, m->dst_stride, m->d_dst_base, slab, &m->gcol, m->d_moves, m->nmoves,
                            m->d_colb, m->ncolb, m->final_shift, m->normalise, stream) != 0) return MPIRFFT_ENODEV;
      if (m->nclass == 2)
      {
         if ((rc = mfft_dsched_run(&m->col2, slab, &m->gcol, m->d_colb2, m->ncolb2, stream)) != 0) return rc;
         if (mfft_dev_finalize(dst, m->dst_stride, m->d_dst_base2, slab, &m->gcol, m->d_moves, m->nmoves,
                               m->d_colb2, m->ncolb2, m->final_shift, m->normalise, stream) != 0) return MPIRFFT_ENODEV;
      }
   }
   return 0;
}

uint64_t mfft_mfa_launches(const mfft_mfa *m)
{
   if (m->fused) return (uint64_t) m->pcol.npasses + m->prow.npasses + m->pcol2.npasses;
   return (uint64_t) m->col.s->nstages + m->row.s->nstages + 1 + (m->nclass == 2 ? m->col2.s->nstages + (m->inverse ? 1 : 0) : 0);
}

/* developer aid (no device needed): the pass structure of a fused MFA plan on stdout --
 * per pass and local stage the number of ops and how many of them are chunk-aligned
 * (all exponents multiples of `align_bits`) for batch column `col` */
void mfft_mfa_debug_dump(int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc, uint32_t col, uint32_t align_bits)
{
   mfft_mfa mm; int which; uint32_t i, k, st;
   if (mfft_mfa_plan(&mm, inverse, n, w, n1, trunc, 0, 0, 0) != 0) { printf("plan failed\n"); return; }
   if (!mm.fused) { printf("not fused\n"); mfft_mfa_free(&mm); return; }
   for (which = 0; which < 2; which++)
   {
      const mfft_passes *P = which ? &mm.prow : &mm.pcol;
      uint64_t M2 = 2*n*w;
      printf("%s: %u passes, batch %u\n", which ? "rows" : "cols", P->npasses, which ? mm.nrowb : mm.ncolb);
      for (i = 0; i < P->npasses; i++)
      {
         const mfft_pass *p = &P->pass[i];
         uint32_t cnt[64] = {0}, al[64] = {0}, two[64] = {0}, ld = 0, stc = 0;
         for (k = 0; k < p->npos_total; k++) { ld += (p->pos[k] & MFFT_TILE_LOAD) ? 1 : 0; stc += (p->pos[k] & MFFT_TILE_STORE) ? 1 : 0; }
         uint32_t tix, r4[64] = {0};
         for (tix = 0; tix < p->ntiles; tix++)
         for (k = p->tiles[tix].op_off; k < p->tiles[tix].op_off + p->tiles[tix].nops; k++)
         {
            const mfft_tileop *o = &p->ops[k];
            if (o->kind == MFFT_K_FWD4 || o->kind == MFFT_K_INV4) { r4[o->lstage < 63 ? o->lstage : 63]++; continue; }
            uint64_t e[4] = { (o->eSA + (uint64_t) col*o->cSA) % M2, (o->eSB + (uint64_t) col*o->cSB) % M2,
                              (o->eTA + (uint64_t) col*o->cTA) % M2, (o->eTB + (uint64_t) col*o->cTB) % M2 };
            int s[4] = { o->sSA, o->b != 0xFFFF ? o->sSB : 0, o->t != 0xFFFF ? o->sTA : 0, (o->t != 0xFFFF && o->b != 0xFFFF) ? o->sTB : 0 };
            int a = 1, j;
            for (j = 0; j < 4; j++) if (s[j] && (e[j] % align_bits)) a = 0;
            st = o->lstage < 63 ? o->lstage : 63;
            cnt[st]++; al[st] += a; two[st] += (o->t != 0xFFFF);
         }
         printf("  pass %u: tiles %u, max_npos %u, stages %u, ops %u (radix-4 units %u, run-time decoded %u), loads %u, stores %u\n", i, p->ntiles, p->max_npos, p->nstages, p->nops_total, p->nr4, p->nany, ld, stc);
         for (st = 0; st < p->nstages && st < 64; st++) printf("     stage %2u: radix-2 ops %6u  aligned %6u  two-output %6u   radix-4 units %6u\n", st, cnt[st], al[st], two[st], r4[st]);
      }
   }
   mfft_mfa_free(&mm);
}

/* developer aid: the ops of one tile of one pass (which: 0 cols, 1 rows) */
void mfft_mfa_debug_tile(int inverse, uint64_t n, uint64_t w, uint64_t n1, uint64_t trunc, int which, uint32_t pass, uint32_t tile)
{
   mfft_mfa mm; uint32_t k;
   if (mfft_mfa_plan(&mm, inverse, n, w, n1, trunc, 0, 0, 0) != 0 || !mm.fused) { printf("plan failed\n"); return; }
   {
      const mfft_passes *P = which ? &mm.prow : &mm.pcol;
      const mfft_pass *p = &P->pass[pass]; const mfft_tile *t = &p->tiles[tile];
      for (k = t->op_off; k < t->op_off + t->nops; k++)
      {
         const mfft_tileop *o = &p->ops[k];
         printf("st %u kind %u a %u b %u s %u t %u kparam %08x | e %u %u %u %u s %d %d %d %d\n", o->lstage, o->kind, o->a, o->b, o->s, o->t,
                o->kparam, o->eSA, o->eSB, o->eTA, o->eTB, o->sSA, o->sSB, o->sTA, o->sTB);
      }
   }
   mfft_mfa_free(&mm);
}
