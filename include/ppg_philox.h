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

#endif /* PPG_PHILOX_H_ */
