// ppg_obs.cu — the observation writer of the two-kernel step (all env variants).
//
// `_get_observation` (BASE:511-539, ECO:700-730, STAG:944-1008) produces > 90 % of the bytes of a step and has no
// order dependence at all, while the rest of `step()` is a latency-bound walk through order-dependent phases.  The
// step kernels (ppg_base.cu / ppg_eco.cu / ppg_stag.cu) therefore stop after the state update and leave, per env, an
// IMAGE in HBM: [header | fp32 value tables | row descriptors | padded owner maps | wall table] — exactly the bytes
// their in-kernel row writer would have gathered from.  This kernel turns images into observation rows:
//
//   * one CTA (4 warps) works on one env at a time; envs are drawn from a ticket counter (populations differ 10x);
//   * the image (5–10 KB) arrives in shared memory as ONE bulk asynchronous copy (cp.async.bulk global -> shared,
//     TMA engine, completion on an mbarrier), double buffered: the copy of the next env is in flight while the rows
//     of the current one are written;
//   * a warp writes whole rows: per lane 8–13 map-entry loads, then 8–13 value-table loads at per-lane constant
//     offsets (emit_row_t, ppg_step_common.cuh), then 2–3 `st.global.cs.v4.f32` — every row is a contiguous 784 /
//     1296 / 1620-byte streaming store;
//   * it also PLACES the newborn rows: by now every env has published its birth count, so the exclusive prefix over
//     the envs before is a plain read (the one-kernel design had every env with newborns spin-wait for its
//     predecessors; that wait was 60 % of the ECO / STAG step kernels, profiles/r01_fused_*).  Row labels of the
//     newborns, `new_off` and the newborns' `ag_prow` (the row their next action is read from) are written here.
//
// Rows of agents that died mid-step were captured by the step kernel at that moment (DSC_SKIP).
#include <cuda_runtime.h>

#include <cstdlib>

#include "ppg_step_common.cuh"

namespace ppg {


__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// bulk asynchronous global -> shared copy (TMA engine, SASS UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(unsigned sdst, const void* gsrc, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
               "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}

#define OBS_MAX_DEFER 256

// The per-lane gather constants of ONE species are kept by a warp ACROSS envs: loading them from global memory cost ~60
// instructions and a global round trip per species and env (13 % of the kernel's instructions when every warp reloaded both
// species for every env, profiles/r02_summary.md).  The CTA holds both species' tables in shared memory (6.6 KB, loaded
// once); a warp reloads from there only when it changes species and re-bases the table addresses when the image buffer
// alternates (nj integer adds).
struct RelCache {
  RowRel rr;
  int s;          // species the constants belong to (-1: none)
  unsigned base;  // virtual image base the table addresses were built for
};
struct RelSmem {
  int2 rel[2][PPG_MAX_NJ][32];
  unsigned self[2][32];
  unsigned short ij[2][PPG_MAX_NJ][32];  // window row | window column << 8 of the element (STAG: cut-off windows)
};

// STAG: a row whose window is cut off (saturated forward view, STAG:944-1008): element (c, i, j) is zero unless i <= ihi and
// j <= jhi.  Straight-line like emit_row_t, the gather constants from the warp's registers, the element's window
// coordinates from shared memory (the generic emit_row_masked reads its constants from global memory element by element).
template <typename MapT, int N, bool VEC>
__device__ __forceinline__ void emit_row_cut_t(const StepParams& p, const RelSmem& t, unsigned sb32, float* dst, int cellp, int s,
                                               const RowRel& r, int ihi, int jhi, int lane) {
  const unsigned a0 = sb32 + (unsigned)(p.so_map[0] + cellp * (int)sizeof(MapT));
  unsigned idx[N];
  float val[N];
#pragma unroll
  for (int j = 0; j < N; ++j) idx[j] = lds_map<MapT>(a0 + (unsigned)r.relb[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) val[j] = lds_f32(r.tbl[j] + 4u * idx[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const unsigned m = t.ij[s][j][lane];
    if ((int)(m & 0xFFu) > ihi || (int)(m >> 8) > jhi) val[j] = 0.f;
  }
  const int elems = p.elems[s];
  if (VEC) {
    float4* d = reinterpret_cast<float4*>(dst) + lane;
#pragma unroll
    for (int v = 0; v < N / 4; ++v)
      if (4 * (lane + 32 * v) < elems) __stcs(d + 32 * v, make_float4(val[4 * v], val[4 * v + 1], val[4 * v + 2], val[4 * v + 3]));
  } else {
    float* d = dst + lane;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (lane + 32 * j < elems) __stcs(d + 32 * j, val[j]);
  }
}
template <typename MapT>
__device__ __forceinline__ void emit_row_cut(const StepParams& p, const RelSmem& t, unsigned sb32, float* dst, int cellp, int s,
                                             const RowRel& r, int ihi, int jhi, int lane) {
  switch (p.emit_kind[s]) {
    case 1: emit_row_cut_t<MapT, 8, true>(p, t, sb32, dst, cellp, s, r, ihi, jhi, lane); break;
    case 2: emit_row_cut_t<MapT, 12, true>(p, t, sb32, dst, cellp, s, r, ihi, jhi, lane); break;
    case 3: emit_row_cut_t<MapT, 13, false>(p, t, sb32, dst, cellp, s, r, ihi, jhi, lane); break;
    case 4: emit_row_cut_t<MapT, 5, false>(p, t, sb32, dst, cellp, s, r, ihi, jhi, lane); break;
    case 5: emit_row_cut_t<MapT, 8, false>(p, t, sb32, dst, cellp, s, r, ihi, jhi, lane); break;
    default: emit_row_masked<MapT>(p, sb32, dst, cellp, s, ihi, jhi, lane);
  }
}
__device__ __forceinline__ void ensure_rel(const StepParams& p, const RelSmem& t, RelCache& c, int s, unsigned vb32, int lane) {
  if (c.s != s) {
    const int nj = p.nj[s];
    c.rr.self = t.self[s][lane];
#pragma unroll
    for (int j = 0; j < PPG_MAX_NJ; ++j) {
      c.rr.relb[j] = 0; c.rr.tbl[j] = 0;
      if (j < nj) {
        const int2 v = t.rel[s][j][lane];
        c.rr.relb[j] = v.x; c.rr.tbl[j] = vb32 + (unsigned)v.y;
      }
    }
    c.s = s;
  } else if (c.base != vb32) {
    const unsigned d = vb32 - c.base;
#pragma unroll
    for (int j = 0; j < PPG_MAX_NJ; ++j) c.rr.tbl[j] += d;
  }
  c.base = vb32;
}

// one observation row of an env from its image in shared memory: descriptor k of species s -> global row `row`
template <typename MapT, int KIND>
__device__ __forceinline__ void obs_one_row(const StepParams& p, const RelSmem& rt, const unsigned char* ibp, unsigned vb32, int env, int s, int k,
                                            int row, const RowRel& rr, int lane, unsigned& rowctr) {
  const uint16_t* dsc = reinterpret_cast<const uint16_t*>(ibp + (p.so_dsc[s] - p.so_img));
  const unsigned* dsx = reinterpret_cast<const unsigned*>(ibp + (p.so_dsx[s] - p.so_img));
  const unsigned d = dsc[k];
  if (d == DSC_SKIP) return;  // captured by the step kernel when the agent died
  const int elems = p.elems[s];
  float* dst = p.obs[s] + (size_t)row * elems;
  if (KIND == 1) {
    if (d == DSC_COPY) {  // captured at birth by the step kernel (the episode ended on this step)
      const float* src = p.born_obs[s] + ((size_t)env * PPG_BORN_K + dsx[k]) * elems;
      for (int q = lane; q < elems; q += 32) __stcs(dst + q, __ldcg(src + q));  // L2: written by another SM during this launch
      return;
    }
    emit_row<MapT, false, true>(p, vb32, dst, (int)d, s, rr, rowctr, lane, __uint_as_float(dsx[k]));
  } else if (KIND == 2) {
    if (d == DSC_ZERO) { zero_row(dst, elems, lane); return; }
    const unsigned x = dsx[k];
    const int ih2 = (int)(x & 0xFFu), jh2 = (int)((x >> 8) & 0xFFu);
    if (ih2 >= p.R[s] - 1 && jh2 >= p.R[s] - 1) emit_row<MapT, false, false>(p, vb32, dst, (int)d, s, rr, rowctr, lane);
    else emit_row_cut<MapT>(p, rt, vb32, dst, (int)d, s, rr, ih2, jh2, lane);
  } else {
    emit_row<MapT, false, false>(p, vb32, dst, (int)d, s, rr, rowctr, lane);
  }
}

// Rows of the agents that acted: the warps of the CTA take them in pairs from one shared counter PER SPECIES, so a warp
// that is busy with something else (warp 0: the next env's fetch, the newborn rows) simply takes fewer.  A warp starts
// with the species whose constants it holds and moves to the other one when nothing is left there.
template <typename MapT, int KIND>
__device__ __forceinline__ void obs_old_rows(const StepParams& p, const unsigned char* ibp, unsigned vb32, int env, const int old_base[2],
                                             const int n[2], int* counter, const RelSmem& rt, RelCache& rc, int lane, unsigned& rowctr) {
  // Rows are handed out in PAIRS of the same species (rows 2 q, 2 q + 1 of the species): two plain rows go through the
  // two-row writer, anything else (the odd last row, skipped / zero / copied / cut-off rows) row by row.
  const int first = rc.s;
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {
    const int s = it == 0 ? first : first ^ 1;
    const int n_s = s == 0 ? n[0] : n[1], base = s == 0 ? old_base[0] : old_base[1];
    const int endq = (n_s + 1) >> 1;
    int* ctr = counter + s;
    if (*reinterpret_cast<volatile int*>(ctr) >= endq) continue;  // nothing left of this species
    int q = 0;
    if (lane == 0) q = atom_shared_add(ctr, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= endq) continue;
    ensure_rel(p, rt, rc, s, vb32, lane);
    const RowRel& rr = rc.rr;
    const uint16_t* dsc = reinterpret_cast<const uint16_t*>(ibp + (p.so_dsc[s] - p.so_img));
    const unsigned* dsx = reinterpret_cast<const unsigned*>(ibp + (p.so_dsx[s] - p.so_img));
    do {
      int qn = 0;
      if (lane == 0) qn = atom_shared_add(ctr, 1);  // the next pair's index arrives while this one is being written
      const int ka = 2 * q, kb = ka + 1;
      bool done = false;
      if (kb < n_s) {
        const unsigned d0 = dsc[ka], d1 = dsc[kb];
        bool plain = d0 < DSC_COPY && d1 < DSC_COPY;  // DSC_COPY is the smallest special descriptor
        float sv0 = 0.f, sv1 = 0.f;
        if (KIND == 1) { sv0 = __uint_as_float(dsx[ka]); sv1 = __uint_as_float(dsx[kb]); }
        if (KIND == 2 && plain) {  // STAG: windows cut off by the saturating forward shift take the masked writer
          const unsigned x0 = dsx[ka], x1 = dsx[kb];
          const int full = p.R[s] - 1;
          plain = (int)(x0 & 0xFFu) >= full && (int)((x0 >> 8) & 0xFFu) >= full && (int)(x1 & 0xFFu) >= full && (int)((x1 >> 8) & 0xFFu) >= full;
        }
        if (plain) {
          const int elems = p.elems[s];
          float* dst0 = p.obs[s] + (size_t)(base + ka) * elems;
          done = emit_row2<MapT, KIND == 1>(p, vb32, dst0, dst0 + elems, (int)d0, (int)d1, s, rr, lane, sv0, sv1);
        }
      }
      if (!done) {
        obs_one_row<MapT, KIND>(p, rt, ibp, vb32, env, s, ka, base + ka, rr, lane, rowctr);
        if (kb < n_s) obs_one_row<MapT, KIND>(p, rt, ibp, vb32, env, s, kb, base + kb, rr, lane, rowctr);
      }
      q = __shfl_sync(FULL, qn, 0);
    } while (q < endq);
  }
}

// newborn rows of an env (one warp): descriptors n[s] .. n[s] + births[s] -> rows new_base[s] ..
template <typename MapT, int KIND>
__device__ __forceinline__ void obs_new_rows(const StepParams& p, const unsigned char* ibp, unsigned vb32, int env, const int n[2],
                                             const int births[2], const int new_base[2], const RelSmem& rt, RelCache& rc, int lane, unsigned& rowctr) {
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const int nb = s == 0 ? births[0] : births[1];
    if (nb <= 0) continue;
    const int n_s = s == 0 ? n[0] : n[1], base = s == 0 ? new_base[0] : new_base[1];
    ensure_rel(p, rt, rc, s, vb32, lane);
    for (int j = 0; j < nb; ++j) obs_one_row<MapT, KIND>(p, rt, ibp, vb32, env, s, n_s + j, base + j, rc.rr, lane, rowctr);
  }
}

// labels of the newborn rows, the rows their first actions are read from, new_off (one warp)
__device__ __forceinline__ void obs_newborn_labels(const StepParams& p, int env, const int births[2], const int new_base[2], int lane) {
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const size_t sb = (size_t)env * p.cap[s];
    for (int j = lane; j < births[s]; j += 32) {
      const unsigned long long info = __ldcg(p.nb_info[s] + sb + j);  // L2: written by another SM during this launch
      const int row = new_base[s] + j;
      p.row_env[s][row] = env;
      p.row_agent[s][row] = (int)(info & 0xFFFFu);
      p.reward[s][row] = 0.f;  // a newborn's first reward is 0 in every variant (BASE:443, ECO:1161, STAG:1573)
      p.flags[s][row] = (uint8_t)((info >> 16) & 0xFFu);
      const unsigned dst = (unsigned)(info >> 32) & 0xFFFFu;
      if (dst != 0xFFFFu) p.ag_prow[s][sb + dst] = row;  // the newborn's first action is read from this row
    }
  }
  if (lane < 2) p.new_off[lane][env] = births[lane] > 0 ? new_base[lane] : 0;
}

// KIND: 0 = BASE family, 1 = ECO (own-speed plane), 2 = STAG (cut-off forward view, all-zero rows of ended agents)
//
// Envs arrive through the COMPLETION QUEUE of the step kernel (queue[i] = epoch << 32 | env, pushed after the env's image
// is in HBM), so this kernel may run WHILE the step kernel is still working: it is launched with programmatic stream
// serialization and fills whatever the step kernel's persistent warps leave free on an SM, above all the step
// kernel's long tail.  Correctness never depends on the overlap — launched after the step kernel has ended, every
// queue entry is simply there already.
//
// Inside a CTA: two image buffers; thread 0 is the producer (ticket, queue entry, fences, bulk copy of the NEXT env's
// image — ~2 us of latency per env), warp 0 also owns the newborn rows (they need the births of all envs before this
// one: a prefix over L2-resident counters; if those are not all published yet the env's newborn rows are deferred to
// the end of this CTA's work and the image is fetched again).  The rows of the agents that acted — the bulk — are
// handed out one at a time from a shared counter, so the other warps absorb whatever warp 0 is busy with and all
// warps reach the one CTA barrier per env together (with a static split 24 % of the warp time was spent waiting at
// that barrier for warp 0, profiles/r01_final_summary.md).
template <typename MapT, int KIND, int OBS_WARPS>
// Register budget: sized for 28 warps per SM (72 registers, no spills) — at 32 (64 registers) the kernel spilled ~100 B
// per thread and was 3–5 % slower on every config; 24 is as good, 20 and 16 lose (profiles/r01_final_summary.md)
__global__ void __launch_bounds__(OBS_WARPS * 32, 24 / OBS_WARPS) ppg_obs_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_img[];  // 2 image buffers
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ int s_env[2];
  __shared__ int s_row[2][2];  // next pair of the env in buffer 0 / 1, per species
  __shared__ int s_defer[OBS_MAX_DEFER];
  __shared__ RelSmem s_rel;  // both species' per-lane gather constants
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // PDL chain: whatever follows in the stream with programmatic serialization (the action kernel of the next step, then its
  // step kernel) may become resident in the slots this grid leaves free; it blocks in griddepcontrol.wait until this grid
  // has completed.  All CTAs of this grid are resident (or done) before that can happen, so nothing can starve it.
  allow_dependent_launch();
  const unsigned stride = (unsigned)p.img_stride, img_bytes = (unsigned)p.img_bytes;
  const unsigned img0 = smem_u32(smem_img), bar0 = smem_u32(&s_bar[0]);
  const unsigned epoch = p.epoch;
  const int par = (int)(epoch & 1u);
  const int n_old_total[2] = {p.totals[(par ^ 1) * 4 + 0], p.totals[(par ^ 1) * 4 + 1]};

  // thread 0: ticket -> completion-queue entry -> env -> bulk copy of its image into buffer `st`.
  // Returns the env (or B when the tickets are exhausted); -1 if the entry is not there yet and !block.
  int tk_pending = -1;
  auto fetch = [&](unsigned st, bool block) -> int {
    if (tk_pending < 0) tk_pending = (int)(atomicAdd(p.obs_ticket, 1ULL) - p.obs_ticket_base);
    if (tk_pending >= p.B) return p.B;
    unsigned long long w = ld_volatile(p.queue + tk_pending);
    while ((unsigned)(w >> 32) != epoch) {
      if (!block) return -1;
      __nanosleep(256);
      w = ld_volatile(p.queue + tk_pending);
    }
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");  // the image was written through the generic proxy of another SM
    const int e = (int)(unsigned)w;
    tk_pending = -1;
    mbar_expect_tx(bar0 + 8u * st, img_bytes);
    bulk_load(img0 + st * stride, p.obs_img + (size_t)e * stride, img_bytes, bar0 + 8u * st);
    return e;
  };

  // both species' gather constants -> shared memory; the loads are in flight while thread 0 draws the first ticket and starts
  // the first image copy (they were 6 % of the kernel's stall samples when issued and awaited before that)
  constexpr int REL_N = 2 * PPG_MAX_NJ * 32, REL_IT = (REL_N + OBS_WARPS * 32 - 1) / (OBS_WARPS * 32);
  int2 rel_tmp[REL_IT];
#pragma unroll
  for (int k = 0; k < REL_IT; ++k) {
    const int i = tid + k * OBS_WARPS * 32;
    rel_tmp[k] = i < REL_N ? __ldg(p.obs_rel + i) : make_int2(0, 0);
  }
  const unsigned self_tmp = (tid < 64 && p.obs_self) ? __ldg(p.obs_self + tid) : 0u;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    s_row[0][0] = 0; s_row[0][1] = 0; s_row[1][0] = 0; s_row[1][1] = 0;
    s_env[0] = fetch(0, true);
  }
  {
    int2* flat = &s_rel.rel[0][0][0];
    unsigned short* ijf = &s_rel.ij[0][0][0];
#pragma unroll
    for (int k = 0; k < REL_IT; ++k) {
      const int i = tid + k * OBS_WARPS * 32;
      if (i < REL_N) {
        flat[i] = rel_tmp[k];
        if (KIND == 2) {  // only STAG has cut-off windows
          const int s = i / (PPG_MAX_NJ * 32), j = (i / 32) % PPG_MAX_NJ, l = i & 31;
          const int R = p.R[s], q = p.obs_vec[s] ? 4 * (l + 32 * (j >> 2)) + (j & 3) : l + 32 * j, r = q % (R * R);
          ijf[i] = (unsigned short)((r / R) | ((r % R) << 8));
        }
      }
    }
    if (tid < 64) (&s_rel.self[0][0])[tid] = self_tmp;
  }
  __syncthreads();
  int env = s_env[0];
  unsigned stage = 0, phase = 0;  // bit s of phase: parity the barrier of buffer s completes next
  unsigned rowctr = 0;
  int n_defer = 0;  // warp 0
  RelCache rc;
  rc.s = -1; rc.base = 0u;
  ensure_rel(p, s_rel, rc, warp == 0 ? 0 : 1, img0 - (unsigned)p.so_img, lane);  // warp 0 starts with species 0, the others with species 1

  while (env < p.B) {
    // next env: its image streams into the other buffer (all reads of that buffer ended before the last barrier)
    int nxt = -1;
    if (tid == 0) nxt = fetch(stage ^ 1u, false);
    mbar_wait(bar0 + 8u * stage, (phase >> stage) & 1u);
    phase ^= 1u << stage;

    const unsigned char* ibp = smem_img + (size_t)stage * stride;
    const unsigned vb32 = img0 + stage * stride - (unsigned)p.so_img;  // virtual base: so_* offsets address the image
    const int* ih = reinterpret_cast<const int*>(ibp + (p.so_ihdr - p.so_img));
    const int old_base[2] = {ih[IH_OLD_BASE0], ih[IH_OLD_BASE1]};
    const int n[2] = {ih[IH_N0], ih[IH_N1]};
    const int births[2] = {ih[IH_BIRTHS0], ih[IH_BIRTHS1]};

    if (warp == 0) {
      if (births[0] + births[1] > 0) {
        // first newborn row of this env = old rows of all envs + births of the envs before it
        int nb0 = 0, nb1 = 0;
        bool ready = prefix_before(p.cntB[par], p.sum1[par], p.sum2[par], 2, env, epoch, false, lane, nb0, nb1);
        if (!ready) {
          if (n_defer < OBS_MAX_DEFER) {
            if (lane == 0) s_defer[n_defer] = env;
            ++n_defer;
          } else {
            // list full: wait here (the step kernel never waits for this kernel, so the wait ends)
            ready = prefix_before(p.cntB[par], p.sum1[par], p.sum2[par], 2, env, epoch, true, lane, nb0, nb1);
            if (!ready && lane == 0) atomicOr(p.error, 1u);
          }
        }
        if (ready) {
          const int new_base[2] = {n_old_total[0] + nb0, n_old_total[1] + nb1};
          obs_newborn_labels(p, env, births, new_base, lane);
          obs_new_rows<MapT, KIND>(p, ibp, vb32, env, n, births, new_base, s_rel, rc, lane, rowctr);
        }
      } else if (lane < 2) {
        p.new_off[lane][env] = 0;
      }
    }

    obs_old_rows<MapT, KIND>(p, ibp, vb32, env, old_base, n, &s_row[stage][0], s_rel, rc, lane, rowctr);

    if (tid == 0) {
      if (nxt < 0) nxt = fetch(stage ^ 1u, true);
      s_env[stage ^ 1u] = nxt;
      s_row[stage ^ 1u][0] = 0; s_row[stage ^ 1u][1] = 0;  // nobody is using the other buffer's counters between the barriers
    }
    __syncthreads();  // every read of this buffer is done; the next env is known
    env = s_env[stage ^ 1u];
    stage ^= 1u;
  }

  // PDL contract: a grid launched with programmatic stream serialization whose prerequisite executes
  // griddepcontrol.launch_dependents must itself execute griddepcontrol.wait before it may be considered ordered after that
  // grid.  The tickets are exhausted, so the step kernel has pushed every env and is in its last instructions (header /
  // counter / flag stores after the queue push): waiting here costs nothing and makes "this kernel has finished" imply
  // "the step kernel has finished and its writes are visible" for whatever the stream runs next (the next step kernel,
  // snapshot copies, host reads of env_flags).  A no-op when the kernel was launched without the attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // deferred newborn rows (warp 0): by now (the completion queue is exhausted) every env has published its births or is
  // about to.  No other warp touches the buffers any more.
  if (warp != 0) return;
  for (int i = 0; i < n_defer; ++i) {
    __syncwarp();  // s_defer was written by lane 0; the previous image's reads are done
    env = s_defer[i];
    if (lane == 0) {
      mbar_expect_tx(bar0 + 8u * stage, img_bytes);
      bulk_load(img0 + stage * stride, p.obs_img + (size_t)env * stride, img_bytes, bar0 + 8u * stage);
    }
    int nb0 = 0, nb1 = 0;
    if (!prefix_before(p.cntB[par], p.sum1[par], p.sum2[par], 2, env, epoch, true, lane, nb0, nb1)) {
      if (lane == 0) atomicOr(p.error, 1u);
    }
    mbar_wait(bar0 + 8u * stage, (phase >> stage) & 1u);
    phase ^= 1u << stage;
    const unsigned char* ibp = smem_img + (size_t)stage * stride;
    const unsigned vb32 = img0 + stage * stride - (unsigned)p.so_img;
    const int* ih = reinterpret_cast<const int*>(ibp + (p.so_ihdr - p.so_img));
    const int n[2] = {ih[IH_N0], ih[IH_N1]};
    const int births[2] = {ih[IH_BIRTHS0], ih[IH_BIRTHS1]};
    const int new_base[2] = {n_old_total[0] + nb0, n_old_total[1] + nb1};
    obs_newborn_labels(p, env, births, new_base, lane);
    obs_new_rows<MapT, KIND>(p, ibp, vb32, env, n, births, new_base, s_rel, rc, lane, rowctr);
    stage ^= 1u;
  }
}

template <typename MapT, int KIND, int OBS_WARPS>
static cudaError_t launch_obs_t(const StepParams& p, int n_cta, bool overlap, cudaStream_t stream) {
  const size_t smem = 2 * (size_t)p.img_stride;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(ppg_obs_kernel<MapT, KIND, OBS_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_bytes = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_cta);
  cfg.blockDim = dim3(OBS_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = overlap ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ppg_obs_kernel<MapT, KIND, OBS_WARPS>, p);
}

template <typename MapT, int KIND, int OBS_WARPS>
static cudaError_t occupancy_obs_t(const StepParams& p, int* blocks_per_sm) {
  const size_t smem = 2 * (size_t)p.img_stride;
  cudaError_t e = cudaFuncSetAttribute(ppg_obs_kernel<MapT, KIND, OBS_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ppg_obs_kernel<MapT, KIND, OBS_WARPS>, OBS_WARPS * 32, smem);
}

// Warps per CTA: 4 while eight double-buffered CTAs fit the SM's shared memory (images up to ~12 KB: BASE, ECO, STAG with
// byte maps), else 8 — a big image (16-bit maps, cap_live > 254) halves the resident CTAs, and 8 warps per CTA keep the
// SM's warp count up (STAG cap 160+640: 0.385 -> 0.293 ms per launch; BASE loses 12 % with 8, profiles/r01_final_summary.md)
static inline bool obs_wide(const StepParams& p) {
  if (const char* ev = getenv("PPG_OBS_WARPS")) return atoi(ev) >= 8;
  return 2 * (size_t)p.img_stride * 8 > 200 * 1024;
}

#define PPG_OBS_DISPATCH_W(FN, NW, ...)                                                                            \
  do {                                                                                                             \
    const bool m8 = p.map_bytes == 1;                                                                              \
    if (p.variant == PPG_VARIANT_ECO) return m8 ? FN<uint8_t, 1, NW>(__VA_ARGS__) : FN<uint16_t, 1, NW>(__VA_ARGS__);  \
    if (p.variant == PPG_VARIANT_STAG) return m8 ? FN<uint8_t, 2, NW>(__VA_ARGS__) : FN<uint16_t, 2, NW>(__VA_ARGS__); \
    return m8 ? FN<uint8_t, 0, NW>(__VA_ARGS__) : FN<uint16_t, 0, NW>(__VA_ARGS__);                                \
  } while (0)

cudaError_t launch_obs(const StepParams& p, int n_cta, bool overlap, cudaStream_t stream) {
  if (obs_wide(p)) PPG_OBS_DISPATCH_W(launch_obs_t, 8, p, n_cta, overlap, stream);
  PPG_OBS_DISPATCH_W(launch_obs_t, 4, p, n_cta, overlap, stream);
}
cudaError_t obs_occupancy(const StepParams& p, int* blocks_per_sm) {
  if (obs_wide(p)) PPG_OBS_DISPATCH_W(occupancy_obs_t, 8, p, blocks_per_sm);
  PPG_OBS_DISPATCH_W(occupancy_obs_t, 4, p, blocks_per_sm);
}

}  // namespace ppg
