/* hostcopy.h -- worker threads for copies between pageable host memory and pinned staging (hostcopy.c) */
#ifndef MFFT_HOSTCOPY_H
#define MFFT_HOSTCOPY_H
#include <stddef.h>

typedef void (*mfft_hc_fn)(void *arg, size_t i);

/* number of worker threads (started on first use; MPIRFFT_COPY_THREADS, 0 = none: callers then copy directly) */
int  mfft_hc_threads(void);
/* run fn(arg, i) for i < n on the workers; returns at once.  One job at a time (library lock held). */
void mfft_hc_begin(mfft_hc_fn fn, void *arg, size_t n);
/* the calling thread takes chunks of the current job too, until none is left */
void mfft_hc_help(void);
/* wait until all n chunks of the current job have finished */
void mfft_hc_end(size_t n);
#endif
