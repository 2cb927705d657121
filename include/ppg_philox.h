/*
 * ppg_philox.h — counter-based random stream of the normal (non-replay) mode.
 *
 * Philox4x32-10 (Salmon et al., SC'11) keyed by the handle seed, counter =
 * (env index, episode number, draw index / 4, stream id).  Header-only, plain C, shared by the
 * CUDA kernels and by the CPU oracle so that both produce the same draws when no replay tape is
 * loaded.  The reference draws from numpy's PCG64 (BASE:135,170,764); bit-parity with numpy is
 * provided by the replay tape, not by this generator.
 */
#ifndef PPG_PHILOX_H_
#define PPG_PHILOX_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PPG_HD __host__ __device__ __forceinline__
#else
#define PPG_HD static inline
#endif

/* stream ids */
#define PPG_STREAM_PLACEMENT 0u /* reset(): initial cells */
#define PPG_STREAM_SPAWN 1u     /* spawn fallback cell */
#define PPG_STREAM_TRAIT 2u     /* founder / offspring trait draws (ECO, STAG) */
#define PPG_STREAM_CAPTURE 3u   /* STAG capture success draw */
#define PPG_STREAM_FACING 4u    /* STAG predator facing (founders, newborns) */
#define PPG_STREAM_FOUNDERS 5u  /* trait variants: number of founders of an episode (MR:189-192) */
#define PPG_STREAM_ACTION 7u    /* ppg_random_actions */

typedef struct ppg_u32x4 {
  uint32_t v[4];
} ppg_u32x4;

PPG_HD ppg_u32x4 ppg_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  ppg_u32x4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

/* draw number `idx` of stream `stream` of episode `episode` of env `env` */
PPG_HD uint32_t ppg_draw_u32(uint64_t seed, uint32_t env, uint32_t episode, uint32_t stream,
                             uint32_t idx) {
  ppg_u32x4 r = ppg_philox4x32(env, episode, idx >> 2, stream, (uint32_t)seed,
                               (uint32_t)(seed >> 32));
  return r.v[idx & 3u];
}

/* uniform integer in [0, n) by multiply-high (bias < n / 2^32) */
PPG_HD uint32_t ppg_bounded(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

/*
 * Real-valued draws of the normal mode (ECO/STAG trait draws).  Built only from IEEE-754 basic
 * operations (+ - * / sqrt) so that the CUDA kernels (compiled with --fmad=false) and the CPU oracle
 * (-ffp-contract=off) produce bit-identical values; libm / CUDA math functions are not used.
 * Every call consumes whole Philox counters of its stream: `*ctr` is the stream's draw cursor.
 */
#if defined(__CUDA_ARCH__)
#define PPG_D2U(x) ((uint64_t)__double_as_longlong(x))
#define PPG_U2D(x) (__longlong_as_double((long long)(x)))
#define PPG_SQRT(x) sqrt(x)
#else
#include <math.h>
#include <string.h>
static inline uint64_t PPG_D2U(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double PPG_U2D(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define PPG_SQRT(x) sqrt(x)
#endif

/* 53-bit uniform in [0, 1) from two words (numpy's recipe: (a >> 5) * 2^26 + (b >> 6)) / 2^53) */
PPG_HD double ppg_u01(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

/* natural logarithm of a positive normal double: x = m 2^e, m in [sqrt(1/2), sqrt(2));
 * log m = 2 atanh(s), s = (m - 1) / (m + 1), odd series to s^23 (|s| < 0.172: truncation < 2e-18) */
PPG_HD double ppg_log(double x) {
  uint64_t u = PPG_D2U(x);
  int e = (int)((u >> 52) & 0x7FFu) - 1023;
  double m = PPG_U2D((u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double p = 1.0 / 23.0;
  p = p * z + 1.0 / 21.0;
  p = p * z + 1.0 / 19.0;
  p = p * z + 1.0 / 17.0;
  p = p * z + 1.0 / 15.0;
  p = p * z + 1.0 / 13.0;
  p = p * z + 1.0 / 11.0;
  p = p * z + 1.0 / 9.0;
  p = p * z + 1.0 / 7.0;
  p = p * z + 1.0 / 5.0;
  p = p * z + 1.0 / 3.0;
  p = p * z + 1.0;
  return (double)e * 0.6931471805599453 + 2.0 * s * p;
}

PPG_HD double ppg_draw_u01(uint64_t seed, uint32_t env, uint32_t episode, uint32_t stream, uint32_t* ctr) {
  const ppg_u32x4 r = ppg_philox4x32(env, episode, (*ctr)++, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  return ppg_u01(r.v[0], r.v[1]);
}

/* standard normal by Marsaglia's polar method; one Philox counter per attempt */
PPG_HD double ppg_draw_normal(uint64_t seed, uint32_t env, uint32_t episode, uint32_t stream, uint32_t* ctr) {
  for (;;) {
    const ppg_u32x4 r = ppg_philox4x32(env, episode, (*ctr)++, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double a = 2.0 * ppg_u01(r.v[0], r.v[1]) - 1.0, b = 2.0 * ppg_u01(r.v[2], r.v[3]) - 1.0;
    const double q = a * a + b * b;
    if (q < 1.0 && q > 1e-300) return a * PPG_SQRT(-2.0 * ppg_log(q) / q);
  }
}

#endif /* PPG_PHILOX_H_ */
