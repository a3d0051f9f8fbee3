/* oracle/shim/mpir.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Minimal stand-in for MPIR 2.4.0's public header so that the UNMODIFIED
 * reference translation unit (/root/reference/mul_fft.c, which includes
 * "mpir.h", "gmp-impl.h", "longlong.h" at mul_fft.c:36-38) can be compiled
 * against the GMP 6.3.0 runtime that ships in this image without headers
 * (/usr/lib/x86_64-linux-gnu/libgmp.so.10).  Only the surface the reference
 * touches is declared.  Nothing here is copied from MPIR or GMP; prototypes
 * are restated from the documented GMP ABI.
 */
#ifndef ORACLE_SHIM_MPIR_H
#define ORACLE_SHIM_MPIR_H

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long mp_limb_t;
typedef long          mp_limb_signed_t;
typedef long          mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef mp_limb_t *       mp_ptr;
typedef const mp_limb_t * mp_srcptr;

#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64

typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];

typedef struct {
   mpz_t _mp_seed;
   int   _mp_alg;
   union { void *_mp_lc; } _mp_algdata;
} __gmp_randstate_struct;
typedef __gmp_randstate_struct gmp_randstate_t[1];

/* ---- GMP runtime symbols (exported by libgmp.so.10) ---- */
mp_limb_t __gmpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_add(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t __gmpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t __gmpn_mul(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
void      __gmpn_mul_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
int       __gmpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_neg(mp_ptr, mp_srcptr, mp_size_t);
void      __gmpn_random2(mp_ptr, mp_size_t);

void __gmpz_init(mpz_t);
void __gmpz_clear(mpz_t);
void *__gmpz_realloc(mpz_t, mp_size_t);
void __gmpz_set_ui(mpz_t, unsigned long);
void __gmpz_add(mpz_t, const mpz_t, const mpz_t);
void __gmpz_add_ui(mpz_t, const mpz_t, unsigned long);
void __gmpz_sub(mpz_t, const mpz_t, const mpz_t);
void __gmpz_mul(mpz_t, const mpz_t, const mpz_t);
void __gmpz_mul_2exp(mpz_t, const mpz_t, mp_bitcnt_t);
void __gmpz_mod(mpz_t, const mpz_t, const mpz_t);
int  __gmpz_invert(mpz_t, const mpz_t, const mpz_t);
int  __gmpz_cmp(const mpz_t, const mpz_t);
void __gmpz_urandomb(mpz_t, gmp_randstate_t, mp_bitcnt_t);

void __gmp_randinit_default(gmp_randstate_t);
void __gmp_randclear(gmp_randstate_t);
unsigned long __gmp_urandomm_ui(gmp_randstate_t, unsigned long);
int __gmp_printf(const char *, ...);

#define mpn_add_n   __gmpn_add_n
#define mpn_sub_n   __gmpn_sub_n
#define mpn_add_1   __gmpn_add_1
#define mpn_sub_1   __gmpn_sub_1
#define mpn_add     __gmpn_add
#define mpn_mul     __gmpn_mul
#define mpn_mul_n   __gmpn_mul_n
#define mpn_cmp     __gmpn_cmp
#define mpn_lshift(r, s, n, c) __gmpn_lshift((r), (s), (n), (unsigned int)(c))
#define mpn_rshift(r, s, n, c) __gmpn_rshift((r), (s), (n), (unsigned int)(c))

#define mpz_init     __gmpz_init
#define mpz_clear    __gmpz_clear
#define mpz_realloc  __gmpz_realloc
#define mpz_set_ui   __gmpz_set_ui
#define mpz_add      __gmpz_add
#define mpz_add_ui   __gmpz_add_ui
#define mpz_sub      __gmpz_sub
#define mpz_mul      __gmpz_mul
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_mod      __gmpz_mod
#define mpz_invert   __gmpz_invert
#define mpz_cmp      __gmpz_cmp

#define gmp_randinit_default __gmp_randinit_default
#define gmp_randclear        __gmp_randclear
#define gmp_urandomm_ui      __gmp_urandomm_ui
#define gmp_printf           __gmp_printf

#endif
