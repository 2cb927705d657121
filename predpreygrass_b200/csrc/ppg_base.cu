// ppg_base.cu — the BASE-family environment step as one fused sm_100a kernel.
//
// Reproduces, for B independent env instances in lockstep, `PredPreyGrass.step()` and `reset()` of
//   BASE = predpreygrass/non_evolutionary/base_environment/predpreygrass_rllib_env.py
// and its reward variants (dense_rewards, dense_rewards_additive, sparse_rewards_plus_eating,
// sparse_rewards_plus_kickback; see include/ppg.h PPG_REWARD_*).
//
// Mapping: one warp owns one env for the whole step.  The env's agent lists, grass and an fp32
// copy of the grid live in that warp's slice of shared memory; order-dependent phases (movement in
// action-dict order BASE:259-273, engagements over the sorted `self.agents` BASE:279-380, births
// BASE:389-448) run as warp-uniform sequential loops, everything order-free (decay, regrowth,
// lookups "first prey on my cell", occupancy tests, observation windows, row output) runs
// lane-parallel.  Observation rows are written straight to their final place in the compact
// per-species batch with 16-byte streaming stores; rows of agents that die are written at the
// moment of death because the reference captures them then (BASE:287,327).
//
// Row allocation across envs: old rows of a step are allocated by the previous step (the live
// counts are known then), newborn rows by a single-pass decoupled look-back over CTAs at the end
// of the logic phase — no second kernel, no host round trip, deterministic layout.
#include <cuda_runtime.h>

#include "ppg_device.cuh"

namespace ppg {

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------
// shared-memory view of one env
// ------------------------------------------------------------------------------------------------
struct EnvSmem {
  double* E[2];
  double* E0[2];
  double* gE;
  float* grid;  // [3][CH] padded: predator, prey, grass energies (channels 1..3 of BASE:123)
  uint16_t* id[2];
  uint16_t* pos[2];
  uint16_t* ord[2];  // ord[k] = slot of the k-th agent in engagement order
  uint16_t* rnk[2];  // inverse of ord
  uint16_t* par[2];
  uint16_t* gpos;
  uint8_t* act[2];
  uint8_t* flg[2];
  uint8_t* aux[2];  // kickback count
  uint8_t* gmap;    // cell -> grass index + 1
};

__device__ __forceinline__ EnvSmem carve(unsigned char* base, const StepParams& p) {
  EnvSmem s;
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
  s.gE = reinterpret_cast<double*>(base + p.so_gE);
  s.grid = reinterpret_cast<float*>(base + p.so_grid);
  s.gpos = reinterpret_cast<uint16_t*>(base + p.so_gpos);
  s.gmap = base + p.so_gmap;
  return s;
}

// Observation rows from the PADDED grid.
// The shared-memory grid of an env has a halo of P = max((R-1)/2) zero cells around the G x G
// field, stored with row stride PS = G + P (the P-wide gap after a row doubles as the left halo of
// the next row) and P leading pad cells:  IDX(x, y) = P + (x + P) * PS + y,  CH cells per channel.
// A window element (c, i, j) of an agent at (x, y) is then simply  grid[c-1][IDX(x,y) + (i-off)*PS +
// (j-off)]  with no bounds test; channel 0 ("outside the grid", BASE:522-523) reads a per-CTA
// constant table of the same shape (1 in halo/gap cells, 0 on the field).  Per lane the relative
// offsets of the <= 16 elements it writes are precomputed on the host (obs_rel table).
struct RowRel {
  int rel[4][4];   // float offset from &grid[IDX(x,y)] for float4 group `it`, element k
  unsigned valid;  // bit it: group it*32+lane exists
};

__device__ __forceinline__ RowRel load_rel(const StepParams& p, int s, int lane, int wall_delta) {
  RowRel r;
  r.valid = 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p.obs_rel) + (s * 4 + it) * 32 + lane);
    const int e[4] = {v.x, v.y, v.z, v.w};
    if (v.x != 0x7FFFFFFF) r.valid |= 1u << it;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = e[k] >> 16, sp = (int)(short)(e[k] & 0xFFFF);
      r.rel[it][k] = (c == 0 ? wall_delta : (c - 1) * p.CH) + sp;
    }
  }
  return r;
}

// _get_observation (BASE:511-539): one [C][R][R] fp32 row, warp-cooperative, 16-byte streaming stores
__device__ __forceinline__ void write_row(float* __restrict__ dst, const float* __restrict__ cell0,
                                          const RowRel& r, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (r.valid & (1u << it)) {
      const float4 v = make_float4(cell0[r.rel[it][0]], cell0[r.rel[it][1]], cell0[r.rel[it][2]], cell0[r.rel[it][3]]);
      __stcs(reinterpret_cast<float4*>(dst) + (it * 32 + lane), v);
    }
  }
}

// rare path (rows of agents that die mid-step, BASE:287,327): same, offsets decoded on the fly
__device__ __noinline__ void write_row_slow(float* __restrict__ dst, const float* __restrict__ cell0,
                                            const StepParams& p, int s, int lane, int wall_delta) {
  const RowRel r = load_rel(p, s, lane, wall_delta);
  write_row(dst, cell0, r, lane);
}

// any live agent (either species, newborns included) on cell `pos`?  = `pos in set(agent_positions.values())` (BASE:399,754)
__device__ __forceinline__ bool any_agent_at(const EnvSmem& S, const int nl[2], unsigned pos, int lane) {
  bool hit = false;
#pragma unroll
  for (int s = 0; s < 2; ++s)
    for (int i = lane; i < nl[s]; i += 32) hit |= (S.flg[s][i] & F_ALIVE) && S.pos[s][i] == pos;
  return __any_sync(FULL, hit);
}

__device__ __forceinline__ unsigned long long vload(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void vstore(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

#define DESC(flag, epoch, val) (((unsigned long long)(flag) << 62) | ((unsigned long long)((epoch)&0x3FFFFFFFu) << 32) | (unsigned long long)(unsigned)(val))

// ------------------------------------------------------------------------------------------------
// the step kernel: W warps per CTA, one env per warp
// ------------------------------------------------------------------------------------------------
#define IDX(ps) (PP + (int)(((ps) >> 8) + PP) * PS + (int)((ps)&255u))

template <int W>
__global__ void __launch_bounds__(W * 32) ppg_step_base_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_cnt[W][4];
  __shared__ int s_base[W][4];
  __shared__ int s_incl[4];
  __shared__ unsigned s_ticket;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_ticket = (unsigned)(atomicAdd(p.ticket, 1ULL) - p.ticket_base);
  // per-CTA constant "outside the grid" channel (BASE:522-523) in the padded layout
  float* wall = reinterpret_cast<float*>(smem_raw + (size_t)W * p.smem_per_env);
  for (int i = threadIdx.x; i < p.CH; i += W * 32) wall[i] = __ldg(p.wall_tab + i);
  __syncthreads();
  const unsigned cta = s_ticket;
  const int env = (int)cta * W + warp;
  const bool active = env < p.B;
  const EnvSmem S = carve(smem_raw + (size_t)warp * p.smem_per_env, p);
  const int G = p.G, GG = p.GG, PP = p.P, PS = p.PS, CH = p.CH;
  const int wall_delta = (int)(wall - S.grid);
  const int totals_rd = p.epoch & 1;
  const int mode_r = p.reward_mode;
  const bool dense = mode_r == PPG_REWARD_DENSE || mode_r == PPG_REWARD_DENSE_ADDITIVE;
  const bool kick = mode_r == PPG_REWARD_SPARSE_KICKBACK;

  // per-env registers (warp-uniform)
  int n[2] = {0, 0};        // list length at step start (= old rows)
  int births[2] = {0, 0};
  int old_base[2] = {0, 0};
  int next_live[2] = {0, 0};
  int mode = 0;             // 0 none/idle, 1 reset, 2 step
  EnvHdr h;
  unsigned env_flags = 0;
  unsigned st_starved[2] = {0, 0}, st_eaten = 0, st_grass = 0, st_fallback = 0;
  int cur[2] = {0, 0};
  bool over = false, trunc = false;

  if (active) {
    h = p.hdr[env];
    old_base[0] = p.next_off[0][env];
    old_base[1] = p.next_off[1][env];
    if (h.state & ST_NEEDS_RESET) mode = 1;
    else if (h.state & ST_IDLE) mode = 0;
    else mode = 2;
  }

  if (mode == 1) {
    // ------------------------------------------------------------------ reset() (BASE:129-217)
    h.episode += 1;
    h.step = 0;
    h.spawn_draws = 0;
    h.status = 0;
    h.sortflag = 0;
    h.first_step = 1;
    h.state = 0;
    const int n_total = p.n_init[0] + p.n_init[1] + p.n_grass;
    // cells in the order predators, prey, grass (BASE:185-187); staged in the (not yet built) grid area
    int* cells = reinterpret_cast<int*>(S.grid);                 // [n_total], n_total <= GG
    unsigned* first = reinterpret_cast<unsigned*>(S.grid) + GG;  // [GG] draw index that claimed the cell
    bool from_tape = false;
    if (p.tape_cells != nullptr) {
      if (h.tape_pos + n_total <= h.tape_end) {
        for (int i = lane; i < n_total; i += 32) cells[i] = p.tape_cells[h.tape_pos + i];
        h.tape_pos += n_total;
        from_tape = true;
      } else {
        h.status |= PPG_STATUS_TAPE_EXHAUSTED;
      }
    }
    if (!from_tape) {
      // Philox rejection draws, accepted in draw order until n_total unique cells (law of BASE:156-177)
      for (int i = lane; i < GG; i += 32) first[i] = 0xFFFFFFFFu;
      __syncwarp();
      int accepted = 0;
      for (unsigned batch = 0; accepted < n_total; ++batch) {
        const unsigned idx0 = batch * 128u + 4u * lane;
        const ppg_u32x4 r = ppg_philox4x32((unsigned)env, h.episode, idx0 >> 2, PPG_STREAM_PLACEMENT,
                                           (unsigned)h.seed_key, (unsigned)(h.seed_key >> 32));
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
    __syncwarp();
    // founders: slots in numeric id order (BASE:143-145,190-200)
    {
      int k0 = 0;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        for (int i = lane; i < p.n_init[s]; i += 32) {
          const int c = cells[k0 + i];
          S.id[s][i] = (uint16_t)i;
          S.pos[s][i] = (uint16_t)(((c / G) << 8) | (c % G));
          S.E[s][i] = p.init_e[s];
          S.flg[s][i] = F_ALIVE;
          if (kick) S.par[s][i] = 0xFFFF;
          S.ord[s][i] = (uint16_t)i;
        }
        k0 += p.n_init[s];
        n[s] = p.n_init[s];
        h.n_list[s] = (unsigned short)p.n_init[s];
        h.next_idx[s] = (unsigned short)p.n_init[s];
      }
      for (int g = lane; g < p.n_grass; g += 32) {
        const int c = cells[k0 + g];
        S.gpos[g] = (uint16_t)(((c / G) << 8) | (c % G));
        S.gE[g] = p.grass_cap;
      }
    }
    __syncwarp();
    // build the grid (BASE:195,200,208)
    for (int i = lane; i < (3 * CH) / 4; i += 32) reinterpret_cast<float4*>(S.grid)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 2; ++s)
      for (int i = lane; i < n[s]; i += 32) S.grid[s * CH + IDX((unsigned)S.pos[s][i])] = (float)p.init_e[s];
    for (int g = lane; g < p.n_grass; g += 32) S.grid[2 * CH + IDX((unsigned)S.gpos[g])] = (float)p.grass_cap;
    __syncwarp();
    cur[0] = n[0]; cur[1] = n[1];
    next_live[0] = n[0]; next_live[1] = n[1];
    env_flags = PPG_ENV_RESET;
  } else if (mode == 2) {
    // ------------------------------------------------------------------ step() (BASE:219-473)
    n[0] = h.n_list[0]; n[1] = h.n_list[1];
    // clear the grid while the loads are in flight
    for (int i = lane; i < (3 * CH) / 4; i += 32) reinterpret_cast<float4*>(S.grid)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < (GG + 3) / 4; i += 32) reinterpret_cast<unsigned*>(S.gmap)[i] = 0u;
    // load the lists (list order = action-dict order = row order of the previous output)
    unsigned bad = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const size_t b = (size_t)env * p.cap[s];
      for (int i = lane; i < n[s]; i += 32) {
        S.id[s][i] = p.ag_id[s][b + i];
        S.pos[s][i] = p.ag_pos[s][b + i];
        const double e = p.ag_e[s][b + i];
        int a = p.actions[s][p.ag_prow[s][b + i]];
        if ((unsigned)a > 8u) { a = 4; bad = PPG_STATUS_BAD_ACTION; }  // reference: KeyError BASE:502
        if (dense) S.E0[s][i] = e;      // energy_before (ADD:256)
        S.E[s][i] = e - p.loss[s];      // Step 1 (BASE:244-250)
        S.act[s][i] = (uint8_t)a;
        S.flg[s][i] = F_ALIVE;
        S.ord[s][i] = (uint16_t)i;
        S.rnk[s][i] = (uint16_t)i;
        if (kick) { S.par[s][i] = p.ag_par[s][b + i]; S.aux[s][i] = 0; }
      }
    }
    h.status |= (unsigned char)__reduce_or_sync(FULL, bad);
    for (int g = lane; g < p.n_grass; g += 32) {
      const size_t b = (size_t)env * p.n_grass;
      S.gpos[g] = p.gr_pos[b + g];
      const double v = p.gr_e[b + g] + p.grass_gain;  // regrowth (BASE:252-256)
      S.gE[g] = v < p.grass_cap ? v : p.grass_cap;
    }
    __syncwarp();
    // rebuild the grid as it stands after Step 1 (equals the persistent grid)
#pragma unroll
    for (int s = 0; s < 2; ++s)
      for (int b0 = 0; b0 < n[s]; b0 += 32) {
        const int i = b0 + lane;
        const bool v = i < n[s];
        const unsigned m = __ballot_sync(FULL, v);
        if (v) {
          const unsigned ps = S.pos[s][i];
          const unsigned grp = __match_any_sync(m, ps);
          // agents sharing a cell: the one latest in dict order wrote last (BASE:247,250)
          if (lane == 31 - __clz(grp)) S.grid[s * CH + IDX(ps)] = (float)S.E[s][i];
        }
        __syncwarp();
      }
    for (int g = lane; g < p.n_grass; g += 32) {
      const unsigned ps = S.gpos[g];
      S.grid[2 * CH + IDX(ps)] = (float)S.gE[g];
      S.gmap[(ps >> 8) * G + (ps & 255)] = (uint8_t)(g + 1);
    }
    __syncwarp();

    // Step 2: movements, sequential in dict order per species (BASE:259-273,495-509).
    // Warp-uniform: every lane replays the same chain, so no intra-warp sync is needed.
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float* gr = S.grid + s * CH;
      for (int j = 0; j < n[s]; ++j) {
        const unsigned ps = S.pos[s][j];
        const int a = S.act[s][j];
        const int x = ps >> 8, y = ps & 255;
        const int ax = (a * 11) >> 5;  // a / 3 for 0 <= a <= 8
        const int nx0 = min(max(x + ax - 1, 0), G - 1), ny0 = min(max(y + (a - 3 * ax) - 1, 0), G - 1);
        const bool blocked = gr[PP + (nx0 + PP) * PS + ny0] > 0.f;  // own-species channel occupied (BASE:506)
        const int nx = blocked ? x : nx0, ny = blocked ? y : ny0;
        gr[PP + (x + PP) * PS + y] = 0.f;                       // BASE:268,272
        gr[PP + (nx + PP) * PS + ny] = (float)S.E[s][j];        // BASE:269,273
        S.pos[s][j] = (uint16_t)((nx << 8) | ny);
      }
    }

    // deferred `self.agents.sort()` of the previous call (BASE:468): engagement order
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (h.sortflag & (1 << s)) {
        const uint16_t* lr = p.lexrank[s];
        for (int i = lane; i < n[s]; i += 32) S.ord[s][i] = __ldg(lr + S.id[s][i]);
        __syncwarp();
        for (int i = lane; i < n[s]; i += 32) {
          const unsigned key = S.ord[s][i];
          int r = 0;
          for (int k = 0; k < n[s]; ++k) r += S.ord[s][k] < key;
          S.rnk[s][i] = (uint16_t)r;
        }
        __syncwarp();
        for (int i = lane; i < n[s]; i += 32) S.ord[s][S.rnk[s][i]] = (uint16_t)i;
        __syncwarp();
      }
    }

    // Step 3a: predators in engagement order (BASE:279-346)
    for (int k = 0; k < n[0]; ++k) {
      const int slot = S.ord[0][k];
      const unsigned ps = S.pos[0][slot];
      const int cell = IDX(ps);
      double e = S.E[0][slot];
      if (e <= 0.0) {  // starved (BASE:284-301): observation as of now
        __syncwarp();
        write_row_slow(p.obs[0] + (size_t)(old_base[0] + k) * p.elems[0], S.grid + cell, p, 0, lane, wall_delta);
        __syncwarp();
        S.grid[cell] = 0.f;
        S.flg[0][slot] = F_DIED;
        st_starved[0]++;
        continue;
      }
      // first prey in agent_positions order (= lowest id) on my cell (BASE:305-312)
      unsigned best = 0xFFFFFFFFu;
      for (int i = lane; i < n[1]; i += 32)
        if ((S.flg[1][i] & F_ALIVE) && S.pos[1][i] == ps) best = min(best, ((unsigned)S.id[1][i] << 16) | (unsigned)i);
      best = __reduce_min_sync(FULL, best);
      if (best != 0xFFFFFFFFu) {
        const int q = best & 0xFFFF;
        e += S.E[1][q];  // BASE:324 (also when the prey's energy is <= 0)
        S.E[0][slot] = e;
        S.grid[cell] = (float)e;  // BASE:325
        S.flg[0][slot] |= F_ATE;
        __syncwarp();
        write_row_slow(p.obs[1] + (size_t)(old_base[1] + S.rnk[1][q]) * p.elems[1], S.grid + cell, p, 1, lane, wall_delta);  // BASE:327
        __syncwarp();
        S.grid[CH + cell] = 0.f;  // BASE:335
        S.flg[1][q] = F_DIED | F_CAUGHT;
        st_eaten++;
      }
    }
    // Step 3b: prey in engagement order (BASE:347-380)
    for (int k = 0; k < n[1]; ++k) {
      const int slot = S.ord[1][k];
      if (!(S.flg[1][slot] & F_ALIVE)) continue;  // caught above (BASE:281)
      const unsigned ps = S.pos[1][slot];
      const int cell = IDX(ps);
      double e = S.E[1][slot];
      if (e <= 0.0) {
        __syncwarp();
        write_row_slow(p.obs[1] + (size_t)(old_base[1] + k) * p.elems[1], S.grid + cell, p, 1, lane, wall_delta);
        __syncwarp();
        S.grid[CH + cell] = 0.f;
        S.flg[1][slot] = F_DIED;
        st_starved[1]++;
        continue;
      }
      const int g = S.gmap[(ps >> 8) * G + (ps & 255)];
      if (g) {  // BASE:351-372 (a patch with energy 0 is still "eaten")
        e += S.gE[g - 1];
        S.E[1][slot] = e;
        S.grid[CH + cell] = (float)e;
        S.grid[2 * CH + cell] = 0.f;
        S.gE[g - 1] = 0.0;
        S.flg[1][slot] |= F_ATE;
        st_grass++;
      }
    }

    // Step 5: births in engagement order, predators then prey (BASE:389-448)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      for (int b0 = 0; b0 < n[s]; b0 += 32) {
        const int k = b0 + lane;
        int slot = 0;
        bool elig = false;
        if (k < n[s]) {
          slot = S.ord[s][k];
          elig = (S.flg[s][slot] & F_ALIVE) && S.E[s][slot] >= p.thr[s];
        }
        unsigned m = __ballot_sync(FULL, elig);
        while (m) {
          const int l = __ffs(m) - 1;
          m &= m - 1;
          const int ps_slot = __shfl_sync(FULL, slot, l);
          if (h.next_idx[s] >= p.n_possible[s]) continue;  // id pool empty (BASE:395,424)
          if (n[s] + births[s] >= p.cap[s]) { h.status |= PPG_STATUS_SLOT_OVERFLOW; continue; }
          const unsigned pp = S.pos[s][ps_slot];
          const int px = pp >> 8, py = pp & 255;
          int nl[2] = {n[0] + births[0], n[1] + births[1]};
          // _find_available_spawn_position (BASE:738-766)
          int sx = -1, sy = -1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cx = px + (c == 0 ? -1 : (c == 1 ? 1 : 0));
            const int cy = py + (c == 2 ? -1 : (c == 3 ? 1 : 0));
            if (sx < 0 && cx >= 0 && cx < G && cy >= 0 && cy < G) {
              if (!any_agent_at(S, nl, (unsigned)((cx << 8) | cy), lane)) { sx = cx; sy = cy; }
            }
          }
          if (sx < 0) {
            st_fallback++;
            if (p.tape_cells != nullptr && h.tape_pos < h.tape_end) {
              const int c = p.tape_cells[h.tape_pos++];  // recorded np.random.randint choice (BASE:764)
              sx = c / G; sy = c % G;
            } else {
              if (p.tape_cells != nullptr) h.status |= PPG_STATUS_TAPE_EXHAUSTED;
              // uniformly random free cell, ascending cell order, Philox draw
              int n_free = 0;
              for (int c0 = 0; c0 < GG; c0 += 32) {
                const int c = c0 + lane;
                bool fr = c < GG;
                if (fr) {
                  const unsigned cp = (unsigned)(((c / G) << 8) | (c % G));
                  for (int s2 = 0; s2 < 2; ++s2)
                    for (int i = 0; i < nl[s2]; ++i) fr &= !((S.flg[s2][i] & F_ALIVE) && S.pos[s2][i] == cp);
                }
                n_free += __popc(__ballot_sync(FULL, fr));
              }
              if (n_free > 0) {
                int kth = (int)ppg_bounded(ppg_draw_u32(h.seed_key, (unsigned)env, h.episode, PPG_STREAM_SPAWN, h.spawn_draws), (unsigned)n_free);
                h.spawn_draws++;
                for (int c0 = 0; c0 < GG && sx < 0; c0 += 32) {
                  const int c = c0 + lane;
                  bool fr = c < GG;
                  if (fr) {
                    const unsigned cp = (unsigned)(((c / G) << 8) | (c % G));
                    for (int s2 = 0; s2 < 2; ++s2)
                      for (int i = 0; i < nl[s2]; ++i) fr &= !((S.flg[s2][i] & F_ALIVE) && S.pos[s2][i] == cp);
                  }
                  const unsigned fm = __ballot_sync(FULL, fr);
                  const int cnt = __popc(fm);
                  if (kth < cnt) {
                    const int c = c0 + (int)__fns(fm, 0, kth + 1);
                    sx = c / G; sy = c % G;
                  } else {
                    kth -= cnt;
                  }
                }
              }
            }
            if (sx < 0) { h.status |= PPG_STATUS_NO_SPAWN_CELL; continue; }  // reference raises here
          }
          const int cs = n[s] + births[s];
          births[s]++;
          const int child_id = h.next_idx[s]++;  // BASE:396-397
          S.id[s][cs] = (uint16_t)child_id;
          S.pos[s][cs] = (uint16_t)((sx << 8) | sy);
          S.E[s][cs] = p.init_e[s];           // BASE:403
          S.flg[s][cs] = F_ALIVE | F_NEWBORN;
          const double pe = S.E[s][ps_slot] - p.init_e[s];  // BASE:404
          S.E[s][ps_slot] = pe;
          S.grid[s * CH + PP + (sx + PP) * PS + sy] = (float)p.init_e[s];  // BASE:405
          S.grid[s * CH + PP + (px + PP) * PS + py] = (float)pe;            // BASE:406
          S.flg[s][ps_slot] |= F_REPRO;
          if (kick) {  // KICK:434-449
            S.aux[s][cs] = 0;
            S.par[s][cs] = S.id[s][ps_slot];
            S.aux[s][ps_slot] = 0;  // rewards[agent] = reproduction_reward overwrites earlier kickbacks (BASE:409)
            const unsigned gp = S.par[s][ps_slot];
            if (gp != 0xFFFFu) {
              int gs = -1;
              for (int i = lane; i < n[s] + births[s]; i += 32)
                if ((S.flg[s][i] & F_ALIVE) && S.id[s][i] == gp) gs = i;
              gs = __reduce_max_sync(FULL, gs);
              if (gs >= 0) S.aux[s][gs]++;
            }
          }
        }
      }
    }
    __syncwarp();

    // counts, termination, truncation (BASE:456-471)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int c = 0;
      for (int i = lane; i < n[s] + births[s]; i += 32) c += (S.flg[s][i] & F_ALIVE) ? 1 : 0;
      cur[s] = __reduce_add_sync(FULL, c);
    }
    h.step += 1;
    const bool all_term = cur[1] <= 0 || cur[0] <= 0;
    trunc = !all_term && h.step >= p.max_steps;  // BASE's extra truncation call (BASE:228-238) folded in
    over = all_term || trunc;
    env_flags = (all_term ? PPG_ENV_TERMINATED : 0) | (trunc ? PPG_ENV_TRUNCATED : 0);
    if (over) {
      if (p.autoreset) { next_live[0] = p.n_init[0]; next_live[1] = p.n_init[1]; }
    } else {
      next_live[0] = cur[0]; next_live[1] = cur[1];
    }
  } else if (active) {
    env_flags = PPG_ENV_IDLE;
  }

  // ---------------------------------------------------------------- row allocation across CTAs
  if (lane == 0) {
    s_cnt[warp][0] = next_live[0]; s_cnt[warp][1] = next_live[1];
    s_cnt[warp][2] = births[0];    s_cnt[warp][3] = births[1];
  }
  __syncthreads();
  if (warp == 0) {
    int agg[4], excl[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      int a = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) a += s_cnt[w][v];
      agg[v] = a;
      excl[v] = 0;
    }
    unsigned long long* my = p.desc + (size_t)cta * 4;
    const int my_agg = lane == 0 ? agg[0] : lane == 1 ? agg[1] : lane == 2 ? agg[2] : agg[3];
    if (cta > 0) {
      if (lane < 4) vstore(my + lane, DESC(1, p.epoch, my_agg));
      unsigned done = 0;
      int posn = (int)cta - 1;
      unsigned spins = 0;
      while (done != 0xF) {
        const int pred = posn - lane;
        unsigned long long w[4];
        bool ok = true;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          w[v] = pred >= 0 ? vload(p.desc + (size_t)pred * 4 + v) : DESC(2, p.epoch, 0);
          const bool valid = ((unsigned)(w[v] >> 32) & 0x3FFFFFFFu) == (p.epoch & 0x3FFFFFFFu) && (w[v] >> 62) != 0;
          ok &= valid || ((done >> v) & 1);
        }
        if (!__all_sync(FULL, ok)) {
          // a predecessor has not published yet; it is resident (tickets are handed out in start
          // order), so this terminates — the cap only protects the box from a wedged launch
          if (++spins > (1u << 22)) { if (lane == 0) atomicOr(p.error, 1u); break; }
          __nanosleep(100);
          continue;
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if ((done >> v) & 1) continue;
          const unsigned pm = __ballot_sync(FULL, (w[v] >> 62) == 2);
          const int fp = pm ? __ffs(pm) - 1 : 32;
          const int contrib = lane <= fp ? (int)(unsigned)w[v] : 0;
          excl[v] += __reduce_add_sync(FULL, contrib);
          if (fp < 32) done |= 1u << v;
        }
        posn -= 32;
      }
    }
    if (lane < 4) {
      const int e = lane == 0 ? excl[0] : lane == 1 ? excl[1] : lane == 2 ? excl[2] : excl[3];
      vstore(my + lane, DESC(2, p.epoch, e + my_agg));
      s_incl[lane] = e + my_agg;
      int run = e;
      for (int w = 0; w < W; ++w) { s_base[w][lane] = run; run += s_cnt[w][lane]; }
    }
  }
  __syncthreads();

  const int n_old_total[2] = {p.next_off[0][p.B + 1 + totals_rd], p.next_off[1][p.B + 1 + totals_rd]};
  if (cta == gridDim.x - 1 && warp == 0 && lane < 2) {
    // last CTA in ticket order: totals of this output and of the next one
    const int s = lane;
    p.n_rows[s] = n_old_total[s];
    p.n_rows[2 + s] = s_incl[2 + s];
    p.old_off[s][p.B] = n_old_total[s];
    p.new_off[s][p.B] = n_old_total[s] + s_incl[2 + s];
    p.next_off[s][p.B + 1 + (totals_rd ^ 1)] = s_incl[s];
  }
  if (!active) return;

  int new_base[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    new_base[s] = n_old_total[s] + s_base[warp][2 + s];
    if (lane == 0) {
      p.old_off[s][env] = old_base[s];
      p.new_off[s][env] = new_base[s];
      p.next_off[s][env] = s_base[warp][s];
    }
  }

  // ------------------------------------------------- rows: metadata, observations, state write-back
  if (mode != 0) {
    const bool keep = !(over && p.autoreset);  // lists of a finished env are dead when it auto-resets
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const RowRel rr = load_rel(p, s, lane, wall_delta);
      const int tot = n[s] + births[s];
      const size_t sb = (size_t)env * p.cap[s];
      float* obs_s = p.obs[s];
      const int elems = p.elems[s];
      int wpos = 0;
      for (int b0 = 0; b0 < tot; b0 += 32) {
        const int k = b0 + lane;
        int cellidx = -1, row = 0, slot = 0;
        if (k < tot) {
          const bool newborn = k >= n[s];
          slot = newborn ? k : S.ord[s][k];
          row = newborn ? new_base[s] + (k - n[s]) : old_base[s] + k;
          const unsigned f = S.flg[s][slot];
          const double e = S.E[s][slot];
          double rew = 0.0;
          if (mode == 2 && !newborn) {
            if (dense) {
              const double e0 = S.E0[s][slot];
              if (f & F_DIED) rew = (f & F_CAUGHT) ? (0.0 - e0) : (e - e0);  // ADD:308,346
              else rew = (e - e0) + ((mode_r == PPG_REWARD_DENSE_ADDITIVE && (f & F_REPRO)) ? p.r_repro[s] : 0.0);  // ADD:468-471
            } else {
              if (f & F_DIED) rew = (f & F_CAUGHT) ? p.pen_caught : 0.0;  // BASE:288,328
              else {
                rew = s == 0 ? ((f & F_ATE) ? p.r_catch : p.r_pstep) : ((f & F_ATE) ? p.r_eat : p.r_qstep);  // BASE:322,341,365,375
                if (f & F_REPRO) rew = p.r_repro[s];  // BASE:409,438 overwrites
                if (kick) for (int q = S.aux[s][slot]; q > 0; --q) rew += p.r_kick[s];  // KICK:446
              }
            }
          }
          unsigned rf = 0;
          if (f & F_DIED) rf |= PPG_ROW_TERMINATED;
          if ((f & F_ALIVE) && trunc) rf |= PPG_ROW_TRUNCATED;
          if (f & F_NEWBORN) rf |= PPG_ROW_NEWBORN;
          if (mode == 1) rf |= PPG_ROW_FOUNDER;
          if (f & F_ATE) rf |= PPG_ROW_ATE;
          p.row_env[s][row] = env;
          p.row_agent[s][row] = S.id[s][slot];
          p.reward[s][row] = (float)rew;
          p.flags[s][row] = (uint8_t)rf;
          if (f & F_ALIVE) cellidx = IDX((unsigned)S.pos[s][slot]);
        }
        unsigned m = __ballot_sync(FULL, cellidx >= 0);
        // survivors in engagement order (= `self.agents` after the sort), then newborns (BASE:398,468)
        if (keep && cellidx >= 0) {
          const int dst = wpos + __popc(m & ((1u << lane) - 1));
          p.ag_id[s][sb + dst] = S.id[s][slot];
          p.ag_pos[s][sb + dst] = S.pos[s][slot];
          p.ag_e[s][sb + dst] = S.E[s][slot];
          p.ag_prow[s][sb + dst] = row;
          if (kick) p.ag_par[s][sb + dst] = S.par[s][slot];
        }
        wpos += __popc(m);
        // Step 6: observations of everyone still present, from the end-of-step grid (BASE:451-453)
        while (m) {
          const int l = __ffs(m) - 1;
          m &= m - 1;
          const int ci = __shfl_sync(FULL, cellidx, l);
          const int r = __shfl_sync(FULL, row, l);
          write_row(obs_s + (size_t)r * elems, S.grid + ci, rr, lane);
        }
      }
      if (keep) h.n_list[s] = (unsigned short)wpos;
    }
    if (over) {
      h.state = p.autoreset ? ST_NEEDS_RESET : ST_IDLE;  // idle: final state stays readable
    }
    if (keep) {
      unsigned char sf = 0;
      if (mode == 2) {
        if (births[0] > 0 || h.first_step) sf |= 1;
        if (births[1] > 0 || h.first_step) sf |= 2;
        h.first_step = 0;
      }
      h.sortflag = sf;
      const size_t gb = (size_t)env * p.n_grass;
      for (int g = lane; g < p.n_grass; g += 32) {
        p.gr_e[gb + g] = S.gE[g];
        if (mode == 1) p.gr_pos[gb + g] = S.gpos[g];
      }
    }
    if (lane == 0) p.hdr[env] = h;
    // per-env counters (PPG_STAT_*)
    if (lane < PPG_N_STATS) {
      unsigned add = 0;
      if (mode == 2) {
        switch (lane) {
          case PPG_STAT_ENV_STEPS: add = 1; break;
          case PPG_STAT_AGENT_STEPS: add = n[0] + n[1]; break;
          case PPG_STAT_EPISODES: add = over; break;
          case PPG_STAT_EPISODE_STEPS: add = over ? h.step : 0; break;
          case PPG_STAT_BIRTHS_PRED: add = births[0]; break;
          case PPG_STAT_BIRTHS_PREY: add = births[1]; break;
          case PPG_STAT_STARVED_PRED: add = st_starved[0]; break;
          case PPG_STAT_STARVED_PREY: add = st_starved[1]; break;
          case PPG_STAT_EATEN_PREY: add = st_eaten; break;
          case PPG_STAT_GRASS_EATEN: add = st_grass; break;
          case PPG_STAT_TRUNCATED: add = trunc; break;
          case PPG_STAT_SPAWN_FALLBACK: add = st_fallback; break;
          default: break;
        }
      }
      if (lane == PPG_STAT_ROWS_PRED) add = n[0] + births[0];
      if (lane == PPG_STAT_ROWS_PREY) add = n[1] + births[1];
      if (add) p.counters[(size_t)env * PPG_N_STATS + lane] += add;
    }
  }
  if (lane == 0) {
    p.env_flags[env] = (uint8_t)env_flags;
    p.env_status[env] = h.status;
    p.env_step[env] = h.step;
    p.env_count[2 * env] = mode == 0 ? h.n_list[0] : cur[0];
    p.env_count[2 * env + 1] = mode == 0 ? h.n_list[1] : cur[1];
  }
}

// ------------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------------

// exclusive scan of the live counts -> first old row of every env in the next output.
// Single CTA; used after ppg_create / ppg_reset / ppg_restore (the step kernel maintains it itself).
__global__ void ppg_prepare_offsets_kernel(const EnvHdr* __restrict__ hdr, int B, int n_init0, int n_init1,
                                           int32_t* off0, int32_t* off1, int totals_slot) {
  __shared__ int s_part[2][1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (B + T - 1) / T;
  const int lo = min(t * per, B), hi = min(lo + per, B);
  int a0 = 0, a1 = 0;
  for (int e = lo; e < hi; ++e) {
    const EnvHdr h = hdr[e];
    const bool rs = h.state & ST_NEEDS_RESET, idle = (h.state & ST_IDLE) && !rs;
    a0 += idle ? 0 : (rs ? n_init0 : h.n_list[0]);
    a1 += idle ? 0 : (rs ? n_init1 : h.n_list[1]);
  }
  s_part[0][t] = a0; s_part[1][t] = a1;
  __syncthreads();
  if (t == 0) {
    int r0 = 0, r1 = 0;
    for (int i = 0; i < T; ++i) { int x0 = s_part[0][i], x1 = s_part[1][i]; s_part[0][i] = r0; s_part[1][i] = r1; r0 += x0; r1 += x1; }
    off0[B + 1 + totals_slot] = r0; off1[B + 1 + totals_slot] = r1;
    off0[B] = r0; off1[B] = r1;
  }
  __syncthreads();
  int r0 = s_part[0][t], r1 = s_part[1][t];
  for (int e = lo; e < hi; ++e) {
    const EnvHdr h = hdr[e];
    const bool rs = h.state & ST_NEEDS_RESET, idle = (h.state & ST_IDLE) && !rs;
    off0[e] = r0; off1[e] = r1;
    r0 += idle ? 0 : (rs ? n_init0 : h.n_list[0]);
    r1 += idle ? 0 : (rs ? n_init1 : h.n_list[1]);
  }
}

// after ppg_restore: define the "previous output" of the restored state as each env's list laid out
// densely in list order, so that actions for the next step can be indexed by row again
__global__ void ppg_relabel_rows_kernel(StepParams p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.B) return;
  const EnvHdr h = p.hdr[e];
  const bool live = !(h.state & (ST_NEEDS_RESET | ST_IDLE));
  for (int s = 0; s < 2; ++s) {
    const int base = p.next_off[s][e];
    p.old_off[s][e] = base;
    p.new_off[s][e] = p.next_off[s][p.B];
    if (e == 0) {
      p.old_off[s][p.B] = p.next_off[s][p.B];
      p.new_off[s][p.B] = p.next_off[s][p.B];
      p.n_rows[s] = p.next_off[s][p.B];
      p.n_rows[2 + s] = 0;
    }
    if (!live) continue;
    const size_t b = (size_t)e * p.cap[s];
    for (int j = 0; j < h.n_list[s]; ++j) {
      p.ag_prow[s][b + j] = base + j;
      p.row_env[s][base + j] = e;
      p.row_agent[s][base + j] = p.ag_id[s][b + j];
      p.reward[s][base + j] = 0.f;
      p.flags[s][base + j] = 0;
    }
  }
  p.env_flags[e] = 0;
  p.env_status[e] = h.status;
  p.env_step[e] = h.step;
  p.env_count[2 * e] = h.n_list[0];
  p.env_count[2 * e + 1] = h.n_list[1];
}

__global__ void ppg_init_hdr_kernel(EnvHdr* hdr, int B, unsigned long long seed) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  EnvHdr h = {};
  h.seed_key = seed;
  h.state = ST_IDLE;  // not reset yet
  hdr[e] = h;
}

// schedule reset() for the masked envs (mask NULL = all), optionally re-keying the Philox stream
__global__ void ppg_mark_reset_kernel(EnvHdr* hdr, int B, const unsigned long long* seeds, const uint8_t* mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  if (mask && !mask[e]) return;
  if (seeds) hdr[e].seed_key = seeds[e];
  hdr[e].state = ST_NEEDS_RESET;
}

__global__ void ppg_set_tape_kernel(EnvHdr* hdr, int B, const long long* cell_off) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  hdr[e].tape_pos = cell_off ? cell_off[e] : 0;
  hdr[e].tape_end = cell_off ? cell_off[e + 1] : 0;
}

// uniform random actions for the rows of the last output (synthetic rollouts)
__global__ void ppg_random_actions_kernel(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ row_env0,
                                          const int32_t* __restrict__ row_agent0, const int32_t* __restrict__ row_env1,
                                          const int32_t* __restrict__ row_agent1, int32_t* act0, int32_t* act1,
                                          unsigned long long seed, unsigned call, unsigned n_actions) {
  const int n0 = n_rows[0] + n_rows[2], n1 = n_rows[1] + n_rows[3];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    const int s = i >= n0;
    const int row = s ? i - n0 : i;
    const unsigned env = (unsigned)(s ? row_env1[row] : row_env0[row]);
    const unsigned id = (unsigned)(s ? row_agent1[row] : row_agent0[row]);
    const unsigned r = ppg_draw_u32(seed, env, call, PPG_STREAM_ACTION + 8u * (unsigned)s, id);
    (s ? act1 : act0)[row] = (int32_t)ppg_bounded(r, n_actions);
  }
}

// sum the per-env counters into int64 totals (block reduce + one atomic per block and counter)
__global__ void ppg_stats_kernel(const uint32_t* __restrict__ counters, const EnvHdr* __restrict__ hdr, int B,
                                 unsigned long long* out) {
  __shared__ unsigned long long s_acc[PPG_N_STATS];
  if (threadIdx.x < PPG_N_STATS) s_acc[threadIdx.x] = 0;
  __syncthreads();
  const int k = threadIdx.x & (PPG_N_STATS - 1);
  unsigned long long a = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)B * PPG_N_STATS; i += (size_t)gridDim.x * blockDim.x) {
    unsigned v = counters[i];
    if (k == PPG_STAT_STATUS_ENVS) v = hdr[i / PPG_N_STATS].status != 0;
    a += v;
  }
  atomicAdd(&s_acc[k], a);
  __syncthreads();
  if (threadIdx.x < PPG_N_STATS && s_acc[threadIdx.x]) atomicAdd(out + threadIdx.x, s_acc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// launch wrappers used by ppg_api.cu
// ------------------------------------------------------------------------------------------------
template <int W>
static cudaError_t launch_w(const StepParams& p, int n_cta, size_t smem, cudaStream_t stream) {
  static size_t attr_bytes = 0;  // opt in to > 48 KB of dynamic shared memory (grows monotonically)
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(ppg_step_base_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_bytes = smem;
  }
  ppg_step_base_kernel<W><<<n_cta, W * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_step_base(const StepParams& p, int warps_per_cta, int n_cta, size_t smem, cudaStream_t stream) {
  switch (warps_per_cta) {
    case 1: return launch_w<1>(p, n_cta, smem, stream);
    case 2: return launch_w<2>(p, n_cta, smem, stream);
    case 4: return launch_w<4>(p, n_cta, smem, stream);
    case 8: return launch_w<8>(p, n_cta, smem, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_prepare_offsets(const EnvHdr* hdr, int B, int n0, int n1, int32_t* off0, int32_t* off1, int slot, cudaStream_t s) {
  ppg_prepare_offsets_kernel<<<1, 1024, 0, s>>>(hdr, B, n0, n1, off0, off1, slot);
  return cudaGetLastError();
}
cudaError_t launch_relabel_rows(const StepParams& p, cudaStream_t s) {
  ppg_relabel_rows_kernel<<<(p.B + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_init_hdr(EnvHdr* hdr, int B, unsigned long long seed, cudaStream_t s) {
  ppg_init_hdr_kernel<<<(B + 255) / 256, 256, 0, s>>>(hdr, B, seed);
  return cudaGetLastError();
}
cudaError_t launch_mark_reset(EnvHdr* hdr, int B, const unsigned long long* seeds, const uint8_t* mask, cudaStream_t s) {
  ppg_mark_reset_kernel<<<(B + 255) / 256, 256, 0, s>>>(hdr, B, seeds, mask);
  return cudaGetLastError();
}
cudaError_t launch_set_tape(EnvHdr* hdr, int B, const long long* cell_off, cudaStream_t s) {
  ppg_set_tape_kernel<<<(B + 255) / 256, 256, 0, s>>>(hdr, B, cell_off);
  return cudaGetLastError();
}
cudaError_t launch_random_actions(const int32_t* n_rows, const int32_t* re0, const int32_t* ra0, const int32_t* re1,
                                  const int32_t* ra1, int32_t* a0, int32_t* a1, unsigned long long seed, unsigned call,
                                  unsigned n_actions, int blocks, cudaStream_t s) {
  ppg_random_actions_kernel<<<blocks, 256, 0, s>>>(n_rows, re0, ra0, re1, ra1, a0, a1, seed, call, n_actions);
  return cudaGetLastError();
}
cudaError_t launch_stats(const uint32_t* counters, const EnvHdr* hdr, int B, unsigned long long* out, cudaStream_t s) {
  ppg_stats_kernel<<<148, 256, 0, s>>>(counters, hdr, B, out);
  return cudaGetLastError();
}

}  // namespace ppg
