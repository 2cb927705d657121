// ppg_step_common.cuh — device code shared by the step kernels of all env variants (ppg_base.cu, ppg_eco.cu):
// the per-env shared-memory view, the observation-row writers (two-level gather: padded byte maps -> fp32
// value tables), the deterministic cross-env row allocation (hierarchical epoch-tagged counts) and the
// Philox placement / free-cell draws.  See DESIGN.md §3.
#pragma once
#include <cuda_runtime.h>

#include "ppg_device.cuh"

namespace ppg {

#define FULL 0xffffffffu

__device__ __forceinline__ unsigned globaltimer_lo() { unsigned v; asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(v)); return v; }
__device__ __forceinline__ unsigned smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }

// ------------------------------------------------------------------------------------------------
// shared-memory view of one env
// ------------------------------------------------------------------------------------------------
template <typename MapT>
struct EnvSmem {
  double* E[2];
  double* E0[2];
  double* gE;
  float* wt;       // wall table: 1.0 at wall_idx, else 0 (constant)
  float* vt[3];    // fp32 value tables: predators [cap0+2], prey [cap1+1], grass [n_grass+1]; entry 0 = 0
  float* stage;    // 2 row buffers of stage_elems floats
  uint8_t* scr;    // [CH] touch counters / predator marks; all zero between uses
  uint16_t* id[2];
  uint16_t* pos[2];
  uint16_t* ord[2];  // ord[k] = slot of the k-th agent in engagement order
  uint16_t* rnk[2];  // inverse of ord
  uint16_t* par[2];
  MapT* map[3];      // padded: [0],[1] owner maps (slot + 1 of the agent the reference grid shows, 0 = empty), [2] grass index + 1
  uint16_t* gpos;
  uint8_t* act[2];
  uint8_t* flg[2];
  uint8_t* aux[2];  // kickback count
  uint8_t* gtag;    // [n_grass] scratch for the prey chunk conflict test
};

template <typename MapT>
__device__ __forceinline__ EnvSmem<MapT> carve(unsigned char* base, const StepParams& p) {
  EnvSmem<MapT> s;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    s.E[k] = reinterpret_cast<double*>(base + p.so_E[k]);
    s.E0[k] = reinterpret_cast<double*>(base + p.so_E0[k]);
    s.id[k] = reinterpret_cast<uint16_t*>(base + p.so_id[k]);
    s.pos[k] = reinterpret_cast<uint16_t*>(base + p.so_pos[k]);
    s.ord[k] = reinterpret_cast<uint16_t*>(base + p.so_ord[k]);
    s.rnk[k] = reinterpret_cast<uint16_t*>(base + p.so_rnk[k]);
    s.par[k] = reinterpret_cast<uint16_t*>(base + p.so_par[k]);
    s.act[k] = base + p.so_act[k];
    s.flg[k] = base + p.so_flg[k];
    s.aux[k] = base + p.so_aux[k];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s.map[k] = reinterpret_cast<MapT*>(base + p.so_map[k]);
    s.vt[k] = reinterpret_cast<float*>(base + p.so_vt[k]);
  }
  s.gE = reinterpret_cast<double*>(base + p.so_gE);
  s.wt = reinterpret_cast<float*>(base + p.so_wt);
  s.stage = reinterpret_cast<float*>(base + p.so_stage);
  s.scr = base + p.so_scr;
  s.gpos = reinterpret_cast<uint16_t*>(base + p.so_gpos);
  s.gtag = base + p.so_gtag;
  return s;
}

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
// bulk asynchronous shared -> global copy (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_store(void* gdst, unsigned ssrc32, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)), "r"(ssrc32), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_volatile(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
#define TAG(epoch, val) (((unsigned long long)(epoch) << 32) | (unsigned long long)(unsigned)(val))

// padded map index of a packed position (x << 8 | y)
#define CELLP(ps) (PP + ((int)((ps) >> 8) + PP) * PS + (int)((ps)&255u))
#define CELLXY(x, y) (PP + ((x) + PP) * PS + (y))

// two-kernel step: row descriptors inside the env image and the image dump
struct RowDesc {
  uint16_t* dsc[2];  // padded cell index of the window centre of the k-th row (k < n: rows of the agents that acted, in
                     // output order; k >= n: newborns in birth order), or DSC_SKIP / DSC_ZERO
  unsigned* dsx[2];  // ECO: own-speed plane value (float bits); STAG: ihi | jhi << 8
};
__device__ __forceinline__ RowDesc carve_desc(unsigned char* base, const StepParams& p) {
  RowDesc d;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    d.dsc[k] = reinterpret_cast<uint16_t*>(base + p.so_dsc[k]);
    d.dsx[k] = reinterpret_cast<unsigned*>(base + p.so_dsx[k]);
  }
  return d;
}

// Writes the image header and copies the image range of the env's shared-memory slice to HBM (coalesced 16-byte
// stores, default caching: the observation kernel reads it back from L2 a few microseconds later).  mode 0 (idle
// env): header only, zero rows.
__device__ __forceinline__ void dump_image(unsigned char* sbase, const StepParams& p, int env, int mode, bool keep, const int old_base[2],
                                           const int n[2], const int births[2], int lane) {
  int* ih = reinterpret_cast<int*>(sbase + p.so_ihdr);
  if (lane < IH_INTS) {
    int v = 0;
    switch (lane) {
      case IH_OLD_BASE0: v = old_base[0]; break;
      case IH_OLD_BASE1: v = old_base[1]; break;
      case IH_N0: v = mode ? n[0] : 0; break;
      case IH_N1: v = mode ? n[1] : 0; break;
      case IH_BIRTHS0: v = mode ? births[0] : 0; break;
      case IH_BIRTHS1: v = mode ? births[1] : 0; break;
      case IH_MODE: v = mode; break;
      case IH_KEEP: v = keep ? 1 : 0; break;
      default: break;
    }
    ih[lane] = v;
  }
  __syncwarp();
  const uint4* src = reinterpret_cast<const uint4*>(sbase + p.so_img);
  uint4* dst = reinterpret_cast<uint4*>(p.obs_img + (size_t)env * p.img_stride);
  const int n16 = mode ? p.img_bytes >> 4 : (IH_INTS * 4) >> 4;
  #pragma unroll 1
  for (int i = lane; i < n16; i += 32) dst[i] = src[i];
  __syncwarp();
  // completion queue: the observation kernel (possibly already running, see ppg_obs.cu) takes the env from here
  if (lane == 0) {
    __threadfence();
    const unsigned long long slot = atomicAdd(p.q_tail, 1ULL) - p.q_base;
    st_volatile(p.queue + slot, TAG(p.epoch, env));
  }
  __syncwarp();
}

// lets a kernel launched with programmatic stream serialization (the observation kernel) start while this one runs
__device__ __forceinline__ void allow_dependent_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// observation rows
// ------------------------------------------------------------------------------------------------
// per-lane gather constants of one species (obs_rel table of the host)
struct RowRel {
  int relb[PPG_MAX_NJ];      // byte offset of the map entry from the agent's own map-0 entry
  unsigned tbl[PPG_MAX_NJ];  // shared address of the value table
  unsigned self;             // bit j: element j is the agent's own trait plane (ECO speed plane), not a gather
};

__device__ __forceinline__ RowRel load_rel(const StepParams& p, int s, unsigned sb32, int lane) {
  RowRel r;
  r.self = p.obs_self ? __ldg(p.obs_self + s * 32 + lane) : 0u;
#pragma unroll
  for (int j = 0; j < PPG_MAX_NJ; ++j) {
    r.relb[j] = 0; r.tbl[j] = 0;
    if (j < p.nj[s]) {
      const int2 v = __ldg(p.obs_rel + (s * PPG_MAX_NJ + j) * 32 + lane);
      r.relb[j] = v.x; r.tbl[j] = sb32 + (unsigned)v.y;  // lanes past the end of the row (last iteration only) get a harmless in-range pair
    }
  }
  return r;
}

// fp32 copies of the energies the observation channels show (float64 state -> float32 row values)
template <typename MapT>
__device__ __forceinline__ void refresh_tables(const EnvSmem<MapT>& S, const StepParams& p, int nt0, int nt1, int lane) {
  #pragma unroll 1
  for (int i = lane; i < nt0; i += 32) S.vt[0][1 + i] = (float)S.E[0][i];
  #pragma unroll 1
  for (int i = lane; i < nt1; i += 32) S.vt[1][1 + i] = (float)S.E[1][i];
  #pragma unroll 1
  for (int g = lane; g < p.n_grass; g += 32) S.vt[2][1 + g] = (float)S.gE[g];
  if (lane == 0) S.vt[0][0] = 0.f;
  if (lane == 1) S.vt[1][0] = 0.f;
  if (lane == 2) S.vt[2][0] = 0.f;
  if (lane == 3) S.vt[0][p.wall_idx] = 0.f;
  __syncwarp();
}

// shared-memory accesses by 32-bit shared address (PTX keeps them in the order written: all map loads of a
// row, then all table loads, then all stores, so the hardware sees 2 dependent latencies per row)
template <typename MapT>
__device__ __forceinline__ unsigned lds_map(unsigned a);
template <>
__device__ __forceinline__ unsigned lds_map<uint8_t>(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
template <>
__device__ __forceinline__ unsigned lds_map<uint16_t>(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// _get_observation (BASE:511-539) of the agent whose padded cell index is `cellp` into global row `dst`.
// Every lane produces N elements of the row, element = table[map[cell + const]]:
//   VEC  (row length a multiple of 4): lane l owns the float4 groups l, l+32, ... (4 consecutive elements each;
//        with byte maps the 32 lanes of one load still hit 32 different banks) and stores them with STG.128;
//   !VEC: lane l owns elements l, l+32, ... and stores 32-bit words, 128 contiguous bytes per warp instruction.
// BULK stages the row in shared memory and hands it to the bulk-copy engine instead (cp.async.bulk, one
// 784/1296/1620-byte copy per row, double buffered); measured slower than direct streaming stores for rows this
// small, kept selectable with PPG_OBS_BULK=1.
template <typename MapT, int N, bool VEC, bool BULK, bool SELF = false>
__device__ __forceinline__ void emit_row_t(const StepParams& p, unsigned sb32, float* dst, int cellp, int s, const RowRel& r,
                                           unsigned& rowctr, int lane, float selfv = 0.f) {
  const unsigned a0 = sb32 + (unsigned)(p.so_map[0] + cellp * (int)sizeof(MapT));
  unsigned idx[N];
  float val[N];
#pragma unroll
  for (int j = 0; j < N; ++j) idx[j] = lds_map<MapT>(a0 + (unsigned)r.relb[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) val[j] = lds_f32(r.tbl[j] + 4u * idx[j]);
  if (SELF) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if ((r.self >> j) & 1u) val[j] = selfv;
  }
  const int elems = p.elems[s];
  if (BULK) {
    const unsigned buf = sb32 + (unsigned)p.so_stage + (rowctr & 1u) * (unsigned)(p.stage_elems * 4);
    ++rowctr;
    if (lane == 0) bulk_wait_read<1>();  // the copy issued two rows ago has finished reading `buf`
    __syncwarp();
    if (VEC) {
#pragma unroll
      for (int v = 0; v < N / 4; ++v)
        if (4 * (lane + 32 * v) < elems) {
#pragma unroll
          for (int k = 0; k < 4; ++k) sts_f32(buf + 16u * (lane + 32 * v) + 4u * k, val[4 * v + k]);
        }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (lane + 32 * j < elems) sts_f32(buf + 4u * (lane + 32 * j), val[j]);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_store(dst, buf, (unsigned)elems * 4u);
      bulk_commit();
    }
  } else {
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
}

// any row shape: one element at a time
template <typename MapT, bool BULK>
__device__ __noinline__ unsigned emit_row_generic(const StepParams& p, unsigned sb32, float* dst, int cellp, int s, unsigned rowctr, int lane,
                                                  unsigned selfmask = 0u, float selfv = 0.f) {
  const unsigned a0 = sb32 + (unsigned)(p.so_map[0] + cellp * (int)sizeof(MapT));
  const bool vec = p.obs_vec[s] != 0;
  unsigned buf = 0;
  if (BULK) {
    buf = sb32 + (unsigned)p.so_stage + (rowctr & 1u) * (unsigned)(p.stage_elems * 4);
    ++rowctr;
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
  }
#pragma unroll 1
  for (int j = 0; j < p.nj[s]; ++j) {
    const int q = vec ? 4 * (lane + 32 * (j >> 2)) + (j & 3) : lane + 32 * j;
    if (q < p.elems[s]) {
      const int2 v = __ldg(p.obs_rel + (s * PPG_MAX_NJ + j) * 32 + lane);
      float x = lds_f32(sb32 + (unsigned)v.y + 4u * lds_map<MapT>(a0 + (unsigned)v.x));
      if ((selfmask >> j) & 1u) x = selfv;
      if (BULK) sts_f32(buf + 4u * q, x); else __stcs(dst + q, x);
    }
  }
  if (BULK) {
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_store(dst, buf, (unsigned)p.elems[s] * 4u);
      bulk_commit();
    }
  }
  return rowctr;
}

// TWO rows of the same species at once (direct stores only): all 2N map loads are issued back to back, then all 2N table
// loads, then the stores — the same two dependent shared-memory latencies as one row, paid once for two rows
// (short_scoreboard was the observation kernel's top stall, profiles/r01_final_summary.md)
template <typename MapT, int N, bool VEC, bool SELF>
__device__ __forceinline__ void emit_row2_t(const StepParams& p, unsigned sb32, float* dst0, float* dst1, int cellp0, int cellp1, int s,
                                            const RowRel& r, int lane, float selfv0, float selfv1) {
  const unsigned a0 = sb32 + (unsigned)(p.so_map[0] + cellp0 * (int)sizeof(MapT));
  const unsigned a1 = sb32 + (unsigned)(p.so_map[0] + cellp1 * (int)sizeof(MapT));
  unsigned idx0[N], idx1[N];
  float val0[N], val1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) idx0[j] = lds_map<MapT>(a0 + (unsigned)r.relb[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) idx1[j] = lds_map<MapT>(a1 + (unsigned)r.relb[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) val0[j] = lds_f32(r.tbl[j] + 4u * idx0[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) val1[j] = lds_f32(r.tbl[j] + 4u * idx1[j]);
  if (SELF) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if ((r.self >> j) & 1u) { val0[j] = selfv0; val1[j] = selfv1; }
  }
  const int elems = p.elems[s];
  if (VEC) {
    float4* d0 = reinterpret_cast<float4*>(dst0) + lane;
    float4* d1 = reinterpret_cast<float4*>(dst1) + lane;
#pragma unroll
    for (int v = 0; v < N / 4; ++v)
      if (4 * (lane + 32 * v) < elems) {
        __stcs(d0 + 32 * v, make_float4(val0[4 * v], val0[4 * v + 1], val0[4 * v + 2], val0[4 * v + 3]));
        __stcs(d1 + 32 * v, make_float4(val1[4 * v], val1[4 * v + 1], val1[4 * v + 2], val1[4 * v + 3]));
      }
  } else {
    float* d0 = dst0 + lane;
    float* d1 = dst1 + lane;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (lane + 32 * j < elems) { __stcs(d0 + 32 * j, val0[j]); __stcs(d1 + 32 * j, val1[j]); }
  }
}

// true if the species' row shape has a two-row writer (else the caller writes the rows one by one)
template <typename MapT, bool SELF>
__device__ __forceinline__ bool emit_row2(const StepParams& p, unsigned sb32, float* dst0, float* dst1, int cellp0, int cellp1, int s,
                                          const RowRel& r, int lane, float selfv0 = 0.f, float selfv1 = 0.f) {
  switch (p.emit_kind[s]) {
    case 1: emit_row2_t<MapT, 8, true, SELF>(p, sb32, dst0, dst1, cellp0, cellp1, s, r, lane, selfv0, selfv1); return true;
    case 2: emit_row2_t<MapT, 12, true, SELF>(p, sb32, dst0, dst1, cellp0, cellp1, s, r, lane, selfv0, selfv1); return true;
    case 3: emit_row2_t<MapT, 13, false, SELF>(p, sb32, dst0, dst1, cellp0, cellp1, s, r, lane, selfv0, selfv1); return true;
    default: return false;
  }
}

template <typename MapT, bool BULK, bool SELF = false>
__device__ __forceinline__ void emit_row(const StepParams& p, unsigned sb32, float* dst, int cellp, int s, const RowRel& r,
                                         unsigned& rowctr, int lane, float selfv = 0.f) {
  switch (p.emit_kind[s]) {  // warp-uniform; the row shapes of the reference's env family get straight-line code
    case 1: emit_row_t<MapT, 8, true, BULK, SELF>(p, sb32, dst, cellp, s, r, rowctr, lane, selfv); break;    // (4,7,7): 49 float4
    case 2: emit_row_t<MapT, 12, true, BULK, SELF>(p, sb32, dst, cellp, s, r, rowctr, lane, selfv); break;   // (4,9,9): 81 float4
    case 3: emit_row_t<MapT, 13, false, BULK, SELF>(p, sb32, dst, cellp, s, r, rowctr, lane, selfv); break;  // (5,9,9): 405 floats
    default: rowctr = emit_row_generic<MapT, BULK>(p, sb32, dst, cellp, s, rowctr, lane, SELF ? r.self : 0u, selfv);
  }
}

// a row whose window is cut off (saturated forward view): element (c, i, j) is zero unless i <= ihi and j <= jhi
template <typename MapT>
__device__ __noinline__ void emit_row_masked(const StepParams& p, unsigned sb32, float* dst, int cellp, int s, int ihi, int jhi, int lane) {
  const unsigned a0 = sb32 + (unsigned)(p.so_map[0] + cellp * (int)sizeof(MapT));
  const bool vec = p.obs_vec[s] != 0;
  const int R = p.R[s], RR = R * R;
#pragma unroll 1
  for (int j = 0; j < p.nj[s]; ++j) {
    const int q = vec ? 4 * (lane + 32 * (j >> 2)) + (j & 3) : lane + 32 * j;
    if (q < p.elems[s]) {
      const int2 v = __ldg(p.obs_rel + (s * PPG_MAX_NJ + j) * 32 + lane);
      float x = lds_f32(sb32 + (unsigned)v.y + 4u * lds_map<MapT>(a0 + (unsigned)v.x));
      const int r = q % RR;
      if (r / R > ihi || r % R > jhi) x = 0.f;
      __stcs(dst + q, x);
    }
  }
}

// ended agents are observed as all-zero rows (STAG:596-612)
__device__ __forceinline__ void zero_row(float* dst, int elems, int lane) {
  #pragma unroll 1
  for (int q = lane; q < elems; q += 32) __stcs(dst + q, 0.f);
}

// rows of agents that die mid-step: the reference captures them at that moment (BASE:287,327)
template <typename MapT, bool BULK, bool SELF = false>
__device__ __noinline__ unsigned emit_row_now(unsigned char* base, const StepParams& p, float* dst, int cellp, int s, int nt0, int nt1,
                                              unsigned rowctr, int lane, float selfv = 0.f) {
  const EnvSmem<MapT> S = carve<MapT>(base, p);
  refresh_tables(S, p, nt0, nt1, lane);
  const unsigned sb32 = (unsigned)__cvta_generic_to_shared(base);
  const RowRel r = load_rel(p, s, sb32, lane);
  emit_row<MapT, BULK, SELF>(p, sb32, dst, cellp, s, r, rowctr, lane, selfv);
  __syncwarp();  // every lane's map / table reads are done before the caller un-writes the dying agent's cell
  return rowctr;
}

// any live agent (either species, newborns included) on cell `pos`?  = `pos in set(agent_positions.values())` (BASE:399,754)
template <typename MapT>
__device__ __forceinline__ bool any_agent_at(const EnvSmem<MapT>& S, const int nl[2], unsigned pos, int lane) {
  bool hit = false;
#pragma unroll
  for (int s = 0; s < 2; ++s)
    #pragma unroll 1
    for (int i = lane; i < nl[s]; i += 32) hit |= (S.flg[s][i] & F_ALIVE) && S.pos[s][i] == pos;
  return __any_sync(FULL, hit);
}

// exclusive prefix, over the envs before `env`, of two 16-bit-packed per-env counts.
//   cnt : per-env words (value pair in bits 31..16 / 15..0), sum1/sum2: per-block / per-group sums at [.][4] + v0, + v0 + 1
// `epoch` = tag the words must carry; wait = poll until they do (this launch's counts) or trust them (previous launch's).
__device__ __forceinline__ bool prefix_before(const unsigned long long* cnt, const unsigned long long* sum1, const unsigned long long* sum2,
                                              int v0, int env, unsigned epoch, bool wait, int lane, int& out0, int& out1) {
  const int blk = env >> 5, grp = env >> 10;
  unsigned spins = 0;
  for (;;) {
    int a0 = 0, a1 = 0;
    bool ok = true;
    if (lane < (env & 31)) {
      const unsigned long long w = ld_volatile(cnt + (blk << 5) + lane);
      ok &= (unsigned)(w >> 32) == epoch;
      a0 += (int)((w >> 16) & 0xFFFFu);
      a1 += (int)(w & 0xFFFFu);
    }
    if (lane < (blk & 31)) {
      const unsigned long long* q = sum1 + (size_t)((grp << 5) + lane) * 4 + v0;
      const unsigned long long w0 = ld_volatile(q), w1 = ld_volatile(q + 1);
      ok &= (unsigned)(w0 >> 32) == epoch && (unsigned)(w1 >> 32) == epoch;
      a0 += (int)(unsigned)w0;
      a1 += (int)(unsigned)w1;
    }
    #pragma unroll 1
    for (int g = lane; g < grp; g += 32) {
      const unsigned long long* q = sum2 + (size_t)g * 4 + v0;
      const unsigned long long w0 = ld_volatile(q), w1 = ld_volatile(q + 1);
      ok &= (unsigned)(w0 >> 32) == epoch && (unsigned)(w1 >> 32) == epoch;
      a0 += (int)(unsigned)w0;
      a1 += (int)(unsigned)w1;
    }
    if (__all_sync(FULL, ok) || !wait) {
      out0 = __reduce_add_sync(FULL, a0);
      out1 = __reduce_add_sync(FULL, a1);
      return __all_sync(FULL, ok);
    }
    // predecessors hold lower tickets, so they are running or done: this terminates.  The cap only
    // protects the box from a wedged launch.
    if (++spins > (1u << 22)) return false;
    __nanosleep(200);
  }
}

// `count` draws of clip(mean + std * N(0,1), lo, hi) into out[0..count), bit-identical to `count` sequential
// ppg_draw_normal calls on the stream (one Philox counter per polar attempt, accepted attempts go to the draws in
// counter order), but with the 32 lanes evaluating 32 consecutive attempts at once.  Returns the advanced counter.
static __device__ __noinline__ unsigned draw_normals_batched(double* out, int count, double mean, double std, double lo, double hi,
                                                             unsigned long long seed_key, unsigned env, unsigned episode, unsigned stream,
                                                             unsigned ctr, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  int got = 0;
  while (got < count) {
    const ppg_u32x4 r = ppg_philox4x32(env, episode, ctr + (unsigned)lane, stream, (unsigned)seed_key, (unsigned)(seed_key >> 32));
    const double a = 2.0 * ppg_u01(r.v[0], r.v[1]) - 1.0, b = 2.0 * ppg_u01(r.v[2], r.v[3]) - 1.0;
    const double q = a * a + b * b;
    const bool acc = q < 1.0 && q > 1e-300;
    const unsigned m = __ballot_sync(FULL, acc);
    const int k = got + __popc(m & lt);
    if (acc && k < count) {
      const double v = mean + std * (a * PPG_SQRT(-2.0 * ppg_log(q) / q));
      out[k] = v < lo ? lo : (v > hi ? hi : v);
    }
    const int cnt = __popc(m);
    if (got + cnt >= count) {
      ctr += __fns(m, 0, count - got) + 1u;  // position after the attempt that produced the last draw
      got = count;
    } else {
      got += cnt;
      ctr += 32u;
    }
  }
  __syncwarp();
  return ctr;
}

// reset(): n_total unique cells in draw order (law of BASE:156-177) from the env's Philox placement stream
static __device__ __noinline__ void philox_placement(int* cells, unsigned* first, int n_total, int GG, unsigned env, unsigned episode,
                                              unsigned long long seed_key, int lane) {
  #pragma unroll 1
  for (int i = lane; i < GG; i += 32) first[i] = 0xFFFFFFFFu;
  __syncwarp();
  int accepted = 0;
  for (unsigned batch = 0; accepted < n_total; ++batch) {
    const unsigned idx0 = batch * 128u + 4u * lane;
    const ppg_u32x4 r = ppg_philox4x32(env, episode, idx0 >> 2, PPG_STREAM_PLACEMENT, (unsigned)seed_key, (unsigned)(seed_key >> 32));
    unsigned cell[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cell[k] = ppg_bounded(r.v[k], (unsigned)GG);
      atomicMin(&first[cell[k]], idx0 + k);
    }
    __syncwarp();
    int mine = 0;
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ok[k] = first[cell[k]] == idx0 + k; mine += ok[k]; }
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    int posn = accepted + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) { if (posn < n_total) cells[posn] = (int)cell[k]; ++posn; }
    accepted += __shfl_sync(FULL, incl, 31);
    __syncwarp();
  }
}

// Occupancy marks for the free-cell scans below: scr[cell] = 1 under every live agent (O(n / 32) per warp instead of
// testing every cell against every agent).  scr is all zero outside its users; unmark_agents restores that.
template <typename MapT>
__device__ __forceinline__ void mark_agents(const EnvSmem<MapT>& S, const StepParams& p, const int nl[2], uint8_t v, int lane) {
  const int PP = p.P, PS = p.PS;
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
    #pragma unroll 1
    for (int i = lane; i < nl[s2]; i += 32)
      if (S.flg[s2][i] & F_ALIVE) S.scr[CELLP((unsigned)S.pos[s2][i])] = v;
  __syncwarp();
}

// number of cells no live agent stands on (`all_positions - occupied_positions`, STAG:1037-1039)
template <typename MapT>
__device__ __noinline__ int count_free_cells(unsigned char* base, const StepParams& p, int nl0, int nl1, int lane) {
  const EnvSmem<MapT> S = carve<MapT>(base, p);
  const int G = p.G, GG = p.GG, PP = p.P, PS = p.PS;
  const int nl[2] = {nl0, nl1};
  mark_agents(S, p, nl, 1, lane);
  int n_free = 0;
  for (int c0 = 0; c0 < GG; c0 += 32) {
    const int c = c0 + lane;
    const bool fr = c < GG && S.scr[CELLXY(c / G, c % G)] == 0;
    n_free += __popc(__ballot_sync(FULL, fr));
  }
  __syncwarp();
  mark_agents(S, p, nl, 0, lane);
  return n_free;
}

// spawn fallback (BASE:760-764): the k-th free cell in ascending cell order, k from the env's Philox spawn stream.
// Returns x << 8 | y, or -1 if no cell is free.
template <typename MapT>
__device__ __noinline__ int philox_free_cell(unsigned char* base, const StepParams& p, int nl0, int nl1, unsigned draw, int lane) {
  const EnvSmem<MapT> S = carve<MapT>(base, p);
  const int G = p.G, GG = p.GG, PP = p.P, PS = p.PS;
  const int nl[2] = {nl0, nl1};
  mark_agents(S, p, nl, 1, lane);
  int n_free = 0;
  for (int c0 = 0; c0 < GG; c0 += 32) {
    const int c = c0 + lane;
    const bool fr = c < GG && S.scr[CELLXY(c / G, c % G)] == 0;
    n_free += __popc(__ballot_sync(FULL, fr));
  }
  int found = -1;
  if (n_free > 0) {
    int kth = (int)ppg_bounded(draw, (unsigned)n_free);
    for (int c0 = 0; c0 < GG; c0 += 32) {
      const int c = c0 + lane;
      const bool fr = c < GG && S.scr[CELLXY(c / G, c % G)] == 0;
      const unsigned fm = __ballot_sync(FULL, fr);
      const int cnt = __popc(fm);
      if (kth < cnt) {
        const int cc = c0 + (int)__fns(fm, 0, kth + 1);
        found = ((cc / G) << 8) | (cc % G);
        break;
      }
      kth -= cnt;
    }
  }
  __syncwarp();
  mark_agents(S, p, nl, 0, lane);
  return found;
}


// Publishes this env's (live, births) counts for the row allocation; the last finisher of each 32-env block /
// 1024-env group publishes the block / group sums, the last group the totals of this output (n_rows).
__device__ __forceinline__ void publish_counts(const StepParams& p, int env, int par, unsigned epoch, const int next_live[2],
                                               const int births[2], int n_blk, int n_grp, const int n_old_total[2], int lane) {
    {
      const int blk = env >> 5, grp = env >> 10;
      bool last = false;
      if (lane == 0) {
        st_volatile(p.cntA[par] + env, TAG(epoch, (next_live[0] << 16) | next_live[1]));
        st_volatile(p.cntB[par] + env, TAG(epoch, (births[0] << 16) | births[1]));
        __threadfence();
        const unsigned bsz = (unsigned)min(32, p.B - (blk << 5));
        last = (atomicAdd(p.done1 + blk, 1u) + 1u) % bsz == 0u;
      }
      if (__shfl_sync(FULL, last, 0)) {  // last env of its 32-env block: block sums
        __threadfence();
        const int e2 = (blk << 5) + lane;
        unsigned long long a = 0, b = 0;
        if (e2 < p.B) { a = ld_volatile(p.cntA[par] + e2); b = ld_volatile(p.cntB[par] + e2); }
        int v[4] = {(int)((a >> 16) & 0xFFFF), (int)(a & 0xFFFF), (int)((b >> 16) & 0xFFFF), (int)(b & 0xFFFF)};
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __reduce_add_sync(FULL, v[q]);
        if (lane < 4) st_volatile(p.sum1[par] + (size_t)blk * 4 + lane, TAG(epoch, lane == 0 ? v[0] : lane == 1 ? v[1] : lane == 2 ? v[2] : v[3]));
        __threadfence();
        __syncwarp();
        bool last2 = false;
        if (lane == 0) {
          const unsigned gsz = (unsigned)min(32, n_blk - (grp << 5));
          last2 = (atomicAdd(p.done2 + grp, 1u) + 1u) % gsz == 0u;
        }
        if (__shfl_sync(FULL, last2, 0)) {  // last block of its group: group sums
          __threadfence();
          const int b2 = (grp << 5) + lane;
          int w[4] = {0, 0, 0, 0};
          if (b2 < n_blk) {
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = (int)(unsigned)ld_volatile(p.sum1[par] + (size_t)b2 * 4 + q);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) w[q] = __reduce_add_sync(FULL, w[q]);
          if (lane < 4) st_volatile(p.sum2[par] + (size_t)grp * 4 + lane, TAG(epoch, lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3]));
          __threadfence();
          __syncwarp();
          bool last3 = false;
          if (lane == 0) last3 = (atomicAdd(p.done3, 1u) + 1u) % (unsigned)n_grp == 0u;
          if (__shfl_sync(FULL, last3, 0)) {  // last group: totals of this output and of the next one
            __threadfence();
            int t[4] = {0, 0, 0, 0};
            #pragma unroll 1
            for (int g = lane; g < n_grp; g += 32) {
#pragma unroll
              for (int q = 0; q < 4; ++q) t[q] += (int)(unsigned)ld_volatile(p.sum2[par] + (size_t)g * 4 + q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) t[q] = __reduce_add_sync(FULL, t[q]);
            if (lane < 2) {
              const int s = lane;
              p.totals[par * 4 + s] = s == 0 ? t[0] : t[1];
              p.totals[par * 4 + 2 + s] = s == 0 ? t[2] : t[3];
              p.n_rows[s] = n_old_total[s];
              p.n_rows[2 + s] = s == 0 ? t[2] : t[3];
              p.old_off[s][p.B] = n_old_total[s];
            }
          }
        }
      }
    }

}

}  // namespace ppg
