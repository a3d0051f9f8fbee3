/* TEST INFRASTRUCTURE: stress of the worker pool of csrc/host/hostcopy.c (built and run by tests/test_hostcopy.py) */
#define _POSIX_C_SOURCE 200809L
#include "hostcopy.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int mfft_dev_bind(void) { return 0; }
typedef struct { unsigned char *cnt; } job;
static void fn(void *a, size_t i) { job *j = (job *) a; __atomic_fetch_add(&j->cnt[i], 1, __ATOMIC_RELAXED); }
int main(void)
{
   int it; size_t i;
   setenv("MPIRFFT_COPY_THREADS", "6", 1);
   printf("threads %d\n", mfft_hc_threads());
   for (it = 0; it < 60000; it++)
   {
      size_t n = 1 + (size_t)(it % 37);
      job j; j.cnt = calloc(n, 1);
      mfft_hc_begin(fn, &j, n);
      if (it & 1) mfft_hc_help();
      mfft_hc_end(n);
      for (i = 0; i < n; i++) if (j.cnt[i] != 1) { printf("BAD it %d i %zu cnt %d\n", it, i, j.cnt[i]); return 1; }
      free(j.cnt);
   }
   printf("ok\n");
   return 0;
}
