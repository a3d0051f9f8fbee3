/* oracle/shim/longlong.h -- TEST INFRASTRUCTURE ONLY.
 * The reference uses exactly two double-limb macros: sub_ddmmss
 * (mul_fft.c:510) and add_ssaaaa (mul_fft.c:3079).  Portable restatements. */
#ifndef ORACLE_SHIM_LONGLONG_H
#define ORACLE_SHIM_LONGLONG_H
#define sub_ddmmss(sh, sl, ah, al, bh, bl) do {                       \
      unsigned __int128 __a = ((unsigned __int128)(ah) << 64) | (al); \
      unsigned __int128 __b = ((unsigned __int128)(bh) << 64) | (bl); \
      __a -= __b; (sl) = (mp_limb_t)__a; (sh) = (mp_limb_t)(__a >> 64); } while (0)
#define add_ssaaaa(sh, sl, ah, al, bh, bl) do {                       \
      unsigned __int128 __a = ((unsigned __int128)(ah) << 64) | (al); \
      unsigned __int128 __b = ((unsigned __int128)(bh) << 64) | (bl); \
      __a += __b; (sl) = (mp_limb_t)__a; (sh) = (mp_limb_t)(__a >> 64); } while (0)
#endif
