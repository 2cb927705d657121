/*
 * ppg_pow.h — bit-exact twin of glibc's double-precision pow() for host and device.
 *
 * The reference computes `speed ** exponent` (ECO:559-563, the locomotion cost) and `(1 - p0) ** max(ratio, 0)`
 * (STAG:1137, the team-capture success probability) with CPython's float power, i.e. libm pow().  glibc's pow (since
 * 2.28: the ARM optimized-routines algorithm, sysdeps/ieee754/dbl-64/e_pow.c) is accurate to ~0.52 ulp but NOT correctly
 * rounded — even `x ** 2.0` differs from the correctly rounded `x * x` by one ulp for ~0.08 % of arguments — so a device
 * result can only be identical to the reference's if it repeats glibc's algorithm step by step: table-driven
 * log(x) = k ln2 + log(c) + log1p(z/c - 1) in double-double, y * log(x) in double-double, table-driven exp with the low
 * part folded in.  This is that algorithm, restated for the variant x86-64 glibc selects at run time on every CPU with
 * FMA (`__pow_fma`, built with __FP_FAST_FMA: the two places below marked FMA use a fused multiply-add, everything else
 * is separately rounded — glibc builds its libm with -ffp-contract=off, the oracle is built the same way and the CUDA
 * library with --fmad=false).  The tables are glibc's own, read from the image's libm.so.6 by
 * scripts/extract_glibc_pow_tables.py (include/ppg_pow_tables.h).
 *
 * Domain: finite x > 0 (normal), finite y — everything the reference feeds it (speeds in [0.5, 2], 1 - p0 in (0, 1),
 * ratio >= 0).  Outside of it (x <= 0, subnormal x, inf / nan) the host version defers to libm and the device version to
 * CUDA's pow(); tests/test_pow_port.py pins the port against the running libm on tens of millions of arguments,
 * including the over/underflow and tiny-|y| branches.
 */
#ifndef PPG_POW_H_
#define PPG_POW_H_

#include <stdint.h>
#include <string.h>

#include "ppg_pow_tables.h"

#if defined(__CUDACC__)
#define PPG_POW_HD __host__ __device__ static inline
#define PPG_POW_CONST __device__ static const
#else
#include <math.h>
#define PPG_POW_HD static inline
#define PPG_POW_CONST static const
#endif

/* tables: on the device they live in global memory (read-only, L1/L2 resident); a second host copy serves host code
 * compiled by nvcc */
#if defined(__CUDACC__)
__device__ static const double ppg_powlog_head_d[9] = PPG_POWLOG_HEAD;
__device__ static const double ppg_powlog_tab_d[128 * 3] = PPG_POWLOG_TAB;
__device__ static const double ppg_exp_head_d[8] = PPG_EXP_HEAD;
__device__ static const unsigned long long ppg_exp_tab_d[256] = PPG_EXP_TAB;
#endif
static const double ppg_powlog_head_h[9] = PPG_POWLOG_HEAD;
static const double ppg_powlog_tab_h[128 * 3] = PPG_POWLOG_TAB;
static const double ppg_exp_head_h[8] = PPG_EXP_HEAD;
static const unsigned long long ppg_exp_tab_h[256] = PPG_EXP_TAB;

#if defined(__CUDA_ARCH__)
#define PPG_POWLOG_HEAD_(i) ppg_powlog_head_d[i]
#define PPG_POWLOG_TAB_(i) ppg_powlog_tab_d[i]
#define PPG_EXP_HEAD_(i) ppg_exp_head_d[i]
#define PPG_EXP_TAB_(i) ppg_exp_tab_d[i]
#define PPG_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define PPG_POWLOG_HEAD_(i) ppg_powlog_head_h[i]
#define PPG_POWLOG_TAB_(i) ppg_powlog_tab_h[i]
#define PPG_EXP_HEAD_(i) ppg_exp_head_h[i]
#define PPG_EXP_TAB_(i) ppg_exp_tab_h[i]
#define PPG_FMA(a, b, c) fma((a), (b), (c))
#endif

PPG_POW_HD uint64_t ppg_asu64(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
PPG_POW_HD double ppg_asf64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

/* log(x) as hi + *tail, x given by its bits (e_pow.c log_inline, __FP_FAST_FMA branch) */
PPG_POW_HD double ppg_pow_log(uint64_t ix, double* tail) {
  const uint64_t OFF = 0x3fe6955500000000ULL;
  const uint64_t tmp = ix - OFF;
  const int i = (int)((tmp >> (52 - 7)) % 128);
  const int k = (int)((int64_t)tmp >> 52); /* arithmetic shift */
  const uint64_t iz = ix - (tmp & (0xfffULL << 52));
  const double z = ppg_asf64(iz);
  const double kd = (double)k;
  const double invc = PPG_POWLOG_TAB_(3 * i), logc = PPG_POWLOG_TAB_(3 * i + 1), logctail = PPG_POWLOG_TAB_(3 * i + 2);
  const double Ln2hi = PPG_POWLOG_HEAD_(0), Ln2lo = PPG_POWLOG_HEAD_(1);
  /* z/c - 1, exact up to the final rounding: FMA */
  const double r = PPG_FMA(z, invc, -1.0);
  /* k*Ln2 + log(c) + r */
  const double t1 = PPG_FMA(kd, Ln2hi, logc);
  const double t2 = t1 + r;
  const double lo1 = PPG_FMA(kd, Ln2lo, logctail);
  const double lo2 = t1 - t2 + r;
  /* A[0] = -0.5 */
  const double ar = PPG_POWLOG_HEAD_(2) * r;
  const double ar2 = r * ar;
  const double ar3 = r * ar2;
  /* k*Ln2 + log(c) + r + A[0]*r*r */
  const double hi = t2 + ar2;
  const double lo3 = PPG_FMA(ar, r, -ar2); /* FMA */
  const double lo4 = t2 - hi + ar2;
  /* p = log1p(r) - r - A[0]*r*r */
  const double A1 = PPG_POWLOG_HEAD_(3), A2 = PPG_POWLOG_HEAD_(4), A3 = PPG_POWLOG_HEAD_(5), A4 = PPG_POWLOG_HEAD_(6),
               A5 = PPG_POWLOG_HEAD_(7), A6 = PPG_POWLOG_HEAD_(8);
  const double q = PPG_FMA(ar2, PPG_FMA(ar2, PPG_FMA(r, A6, A5), PPG_FMA(r, A4, A3)), PPG_FMA(r, A2, A1));
  const double lo = PPG_FMA(ar3, q, lo1 + lo2 + lo3 + lo4); /* lo1 + lo2 + lo3 + lo4 + ar3 * q */
  const double y = hi + lo;
  *tail = hi - y + lo;
  return y;
}

/* e_pow.c specialcase(): the scale 2^(k/N) over- or underflows on its own */
PPG_POW_HD double ppg_pow_exp_special(double tmp, uint64_t sbits, uint64_t ki) {
  double scale, y;
  if ((ki & 0x80000000ULL) == 0) {
    /* k > 0, the exponent of scale might have overflowed by <= 460 */
    sbits -= 1009ULL << 52;
    scale = ppg_asf64(sbits);
    y = 0x1p1009 * PPG_FMA(scale, tmp, scale);
    return y;
  }
  /* k < 0, need special care in the subnormal range */
  sbits += 1022ULL << 52;
  scale = ppg_asf64(sbits);
  const double st = scale * tmp; /* used twice in e_pow.c: the compiler keeps the product, no contraction here */
  y = scale + st;
  const double ay = y < 0.0 ? -y : y;
  if (ay < 1.0) {
    /* round y to the right precision before scaling it into the subnormal range (avoids double rounding) */
    double hi, lo, one = 1.0;
    if (y < 0.0) one = -1.0;
    lo = scale - y + st;
    hi = one + y;
    lo = one - hi + y + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = ppg_asf64(sbits & 0x8000000000000000ULL); /* sign of 0 */
  }
  return 0x1p-1022 * y;
}

/* exp(x + xtail) (e_pow.c exp_inline, sign_bias = 0: the base is positive) */
PPG_POW_HD double ppg_pow_exp(double x, double xtail) {
  uint32_t abstop = (uint32_t)(ppg_asu64(x) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x408u - 0x3c9u) { /* |x| < 2^-54 or |x| >= 512 */
    if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x; /* tiny: WANT_ROUNDING */
    if (abstop >= 0x409u) return (ppg_asu64(x) >> 63) ? 0.0 : ppg_asf64(0x7ff0000000000000ULL); /* under / overflow */
    abstop = 0; /* large x is special-cased below */
  }
  const double InvLn2N = PPG_EXP_HEAD_(0), Shift = PPG_EXP_HEAD_(1), NegLn2hiN = PPG_EXP_HEAD_(2), NegLn2loN = PPG_EXP_HEAD_(3);
  const double C2 = PPG_EXP_HEAD_(4), C3 = PPG_EXP_HEAD_(5), C4 = PPG_EXP_HEAD_(6), C5 = PPG_EXP_HEAD_(7);
  /* exp(x) = 2^(k/N) * exp(r), x = ln2/N*k + r */
  double kd = PPG_FMA(InvLn2N, x, Shift);
  const uint64_t ki = ppg_asu64(kd);
  kd -= Shift;
  double r = PPG_FMA(kd, NegLn2loN, PPG_FMA(kd, NegLn2hiN, x));
  r += xtail;
  const uint64_t idx = 2 * (ki % 128);
  const uint64_t top = ki << (52 - 7);
  const double tail = ppg_asf64(PPG_EXP_TAB_(idx));
  const uint64_t sbits = PPG_EXP_TAB_(idx + 1) + top;
  const double r2 = r * r;
  const double tmp = PPG_FMA(r2 * r2, PPG_FMA(r, C5, C4), PPG_FMA(r2, PPG_FMA(r, C3, C2), tail + r));
  if (abstop == 0) return ppg_pow_exp_special(tmp, sbits, ki);
  const double scale = ppg_asf64(sbits);
  return PPG_FMA(scale, tmp, scale);
}

/* pow(x, y) exactly as glibc computes it (e_pow.c __pow, __FP_FAST_FMA build) */
PPG_POW_HD double ppg_pow(double x, double y) {
  const uint64_t ix = ppg_asu64(x), iy = ppg_asu64(y);
  const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    /* special cases: x < 2^-1022 or negative or inf / nan, |y| < 2^-65 or |y| >= 2^63 or nan */
    const int y_special = 2 * iy - 1 >= 2 * 0x7ff0000000000000ULL - 1;  /* y is 0, inf or nan */
    const int x_ok = topx - 0x001u < 0x7ffu - 0x001u;                    /* finite normal x > 0 */
    if (x_ok && !y_special) {
      /* finite normal x > 0, non-zero finite y with an extreme exponent */
      if (ix == 0x3ff0000000000000ULL) return 1.0;
      if ((topy & 0x7ff) < 0x3be) return ix > 0x3ff0000000000000ULL ? 1.0 + y : 1.0 - y; /* |y| < 2^-65 */
      return ((ix > 0x3ff0000000000000ULL) == (topy < 0x800)) ? ppg_asf64(0x7ff0000000000000ULL) : 0.0;
    }
    if (x_ok && (iy << 1) == 0) return 1.0; /* y == +-0 */
    return pow(x, y);                        /* outside the reference's domain: the platform's pow */
  }
  double lo;
  const double hi = ppg_pow_log(ix, &lo);
  const double ehi = y * hi;
  const double elo = PPG_FMA(y, lo, PPG_FMA(y, hi, -ehi));
  return ppg_pow_exp(ehi, elo);
}

#endif /* PPG_POW_H_ */
