/* oracle/ssmul_oracle.h -- TEST INFRASTRUCTURE ONLY (see ssmul_oracle.c). */
#ifndef SSMUL_ORACLE_H
#define SSMUL_ORACLE_H
#include <stdint.h>
typedef uint64_t limb;
typedef int64_t  slimb;
void orc_normalise(limb *t, long l);
void orc_mul_2exp(limb *r, const limb *a, long l, unsigned long e);
void orc_transform(int kind, limb **ii, long is, long n, unsigned long w, unsigned long ws,
                   unsigned long r, unsigned long c, unsigned long rs, long trunc);
void orc_mfa(int inverse, limb **ii, long n, unsigned long w, long n1, long trunc);
long orc_split_bits(limb **poly, const limb *limbs, long total, long bits, long out);
void orc_combine_bits(limb *res, limb **poly, long length, long bits, long out, long total);
void orc_mulmod(limb *r, const limb *a, const limb *b, long l);
int  orc_new_mpn_mul(limb *r1, const limb *i1, long n1, const limb *i2, long n2, unsigned long depth, unsigned long w);
#endif
