/* hostcopy.c -- operands and results in ordinary (pageable) host memory.
 *
 * A caller that merely swaps libraries hands new_mpn_mul malloc'ed buffers.
 * cudaMemcpyAsync from / to such memory is staged by the driver on the calling thread at 10-13 GB/s
 * and blocks it meanwhile.  Here a few worker threads copy chunks between the caller's buffers and
 * one pinned staging buffer, and the chunks travel by DMA as they become ready: the copy in runs at
 * the workers' combined memcpy rate next to the DMA, the second operand moves while the first one is
 * already being transformed, and the result is copied out chunk by chunk behind the DMA.
 *
 * The pool serves one job at a time; every user holds the library lock (runtime.c), so there is no
 * queue.  Workers never take that lock.
 */
#define _POSIX_C_SOURCE 200809L
#include "hostcopy.h"
#include "runtime.h"
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define HC_MAXW 16

static struct {
   pthread_mutex_t mu; pthread_cond_t cv;
   pthread_t th[HC_MAXW]; int nthreads, started;
   unsigned gen;                                   /* job number, bumped under mu */
   mfft_hc_fn fn; void *arg; size_t n;             /* the job: fn(arg, i) for i < n; stable while a chunk of it is held */
   unsigned long long next;                        /* atomic: job number << 32 | next chunk */
   size_t done;                                    /* atomic: chunks finished */
} hc = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, 0, 0, NULL, NULL, 0, 0, 0 };

/* Take chunks of job `gen` until none is left.  A chunk is claimed by compare-and-swap on (job, index):
   a thread that is late leaving an earlier job can neither claim nor skip a chunk of a later one, and
   whoever holds a claimed chunk keeps its job alive (mfft_hc_end waits for done == n).  When a job has
   ended, `next` holds (job, HC_CLOSED): a late thread that read the count of the FOLLOWING job against an
   index of the ended one fails its compare-and-swap. */
#define HC_CLOSED 0xffffffffu
static void run_chunks(unsigned gen)
{
   for (;;)
   {
      unsigned long long v = __atomic_load_n(&hc.next, __ATOMIC_ACQUIRE);
      if ((unsigned)(v >> 32) != gen || (size_t)(v & 0xffffffffu) >= __atomic_load_n(&hc.n, __ATOMIC_RELAXED)) return;
      if (!__atomic_compare_exchange_n(&hc.next, &v, v + 1, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) continue;
      hc.fn(hc.arg, (size_t)(v & 0xffffffffu));
      __atomic_fetch_add(&hc.done, 1, __ATOMIC_RELEASE);
   }
}

static void *worker(void *unused)
{
   unsigned seen = 0;
   (void) unused;
   for (;;)
   {
      pthread_mutex_lock(&hc.mu);
      while (hc.gen == seen) pthread_cond_wait(&hc.cv, &hc.mu);
      seen = hc.gen;
      pthread_mutex_unlock(&hc.mu);
      mfft_dev_bind();                              /* a worker may wait on an event of the library's device */
      run_chunks(seen);
   }
   return NULL;
}

int mfft_hc_threads(void)
{
   if (!hc.started)
   {
      const char *e = getenv("MPIRFFT_COPY_THREADS");
      long want, ncpu = sysconf(_SC_NPROCESSORS_ONLN);
      int k;
      hc.started = 1;
      want = e ? atol(e) : (ncpu >= 16 ? 6 : ncpu >= 4 ? ncpu / 2 - 1 : 0);
      if (want > HC_MAXW) want = HC_MAXW;
      for (k = 0; k < want; k++)
      {
         pthread_attr_t at;
         pthread_attr_init(&at); pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
         if (pthread_create(&hc.th[hc.nthreads], &at, worker, NULL) == 0) hc.nthreads++;
         pthread_attr_destroy(&at);
      }
   }
   return hc.nthreads;
}

void mfft_hc_begin(mfft_hc_fn fn, void *arg, size_t n)
{
   pthread_mutex_lock(&hc.mu);
   hc.gen++;
   hc.fn = fn; hc.arg = arg;                       /* nobody holds a chunk: the previous job has ended */
   __atomic_store_n(&hc.n, n < HC_CLOSED ? n : HC_CLOSED - 1, __ATOMIC_RELAXED);
   __atomic_store_n(&hc.done, 0, __ATOMIC_RELEASE);
   __atomic_store_n(&hc.next, (unsigned long long) hc.gen << 32, __ATOMIC_RELEASE);
   pthread_cond_broadcast(&hc.cv);
   pthread_mutex_unlock(&hc.mu);
}

void mfft_hc_help(void) { run_chunks(hc.gen); }

void mfft_hc_end(size_t n)
{
   while (__atomic_load_n(&hc.done, __ATOMIC_ACQUIRE) < n) sched_yield();
   __atomic_store_n(&hc.next, ((unsigned long long) hc.gen << 32) | HC_CLOSED, __ATOMIC_RELEASE);
}
