/*
 * ppg_oracle_eco.c — CPU oracle of the ECO env (TEST INFRASTRUCTURE ONLY, see ppg_oracle.h).
 *
 * Sequential restatement of
 *   ECO    = predpreygrass/evolutionary/eco_evolutionary/predpreygrass_rllib_env.py
 *   GENOME = predpreygrass/evolutionary/eco_evolutionary/utils/genome.py
 * with the reference's data structures: dicts keyed by agent id -> per-id arrays, `self.agents`
 * (whose order is also the insertion order of agent_positions / agent_energies: both append at
 * birth and only ever delete), the persistent FLOAT32 grid (ECO:222-224), dead_prey,
 * active_num_* counters (which the reference lets drift, SURVEY §8a quirks 4 and 13 — reproduced
 * literally).  Every block cites the lines it follows.
 *
 * Not restated (host-side analytics that never feed back into the step): agent_event_log,
 * agent_stats_*, per_step_agent_data, lineage tracking (lineage_reward_coeff must be 0, then
 * ECO:943-991 only does `rewards.setdefault(agent, 0.0)`), infos.
 *
 * Pinned against golden trajectories recorded from the unmodified reference
 * (tests/golden/make_golden_eco.py -> tests/golden/eco_*.npz, tests/test_oracle_golden_eco.py).
 *
 * The same file restates the sibling trait variants, which share ECO's skeleton (ppg_config.trait_mode):
 *   MR   = predpreygrass/evolutionary/eco_evolutionary_metabolic_rate/predpreygrass_rllib_env.py
 *   INV  = predpreygrass/evolutionary/eco_evolutionary_investment/predpreygrass_rllib_env.py
 *   COOP = predpreygrass/evolutionary/eco_evolutionary_cooperation/predpreygrass_rllib_env.py
 * (random founder counts, trait-scaled decay / gains, satiation cooldown, density cap, offspring investment, meal
 * sharing; no ageing caps, no carcasses, no speed gating) — each branch cites the variant's lines; pinned by
 * tests/golden/make_golden_traits.py -> tests/golden/{mr,inv,coop}_*.npz, tests/test_oracle_golden_traits.py.
 *   CAD  = predpreygrass/evolutionary/eco_evolutionary_cadence/predpreygrass_rllib_env.py
 * keeps ECO's step (ageing caps, grass intake cap, speed ** exponent in the move cost) and changes what the speed does: it
 * sets the RATE at which a per-agent accumulator lets the agent move (CAD:556-585,674-681), scales the basal cost
 * (CAD:626-633), shows unnormalised in the speed plane (CAD:746) and comes with an action mask (row flag PPG_ROW_FROZEN);
 * predators catch the nearest prey within Chebyshev distance 1 and eat it whole (CAD:837-880), newborns go to a RANDOM
 * free neighbour cell (CAD:795) and start with a random accumulator phase (CAD:1327); no carcasses, no lineage rewards.
 * Pinned by tests/golden/cad_*.npz.
 */
#define IS_MIC(c) ((c)->trait_mode == PPG_TRAIT_METABOLIC || (c)->trait_mode == PPG_TRAIT_INVESTMENT || (c)->trait_mode == PPG_TRAIT_COOPERATION)
#define IS_CAD(c) ((c)->trait_mode == PPG_TRAIT_CADENCE)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ppg_philox.h"
#include "ppg_oracle_int.h"

static inline float* GF(env_t* e, int ch, int x, int y) { return &e->gridf[((size_t)ch * e->G + x) * e->G + y]; }
static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int n_row_channels(const ppg_config* c) { /* + ECO's own-speed plane (ECO:1396-1406) / STAG's visibility channel (STAG:1788) */
  return c->num_obs_channels + (c->include_speed_in_obs ? 1 : 0) + ((c->variant == PPG_VARIANT_STAG && c->include_visibility_channel) ? 1 : 0);
}

void eco_env_alloc(env_t* e) {
  const ppg_config* c = e->c;
  e->gridf = (float*)calloc((size_t)c->num_obs_channels * e->G * e->G, sizeof(float));
  for (int s = 0; s < 2; ++s) {
    size_t n = (size_t)c->n_possible[s];
    e->age[s] = (int32_t*)calloc(n, sizeof(int32_t));
    e->speed[s] = (double*)calloc(n, sizeof(double));
    e->termd[s] = (uint8_t*)calloc(n, 1);
    e->row_elems[s] = n_row_channels(c) * c->obs_range[s] * c->obs_range[s];
  }
  e->max_row_elems = e->row_elems[0] > e->row_elems[1] ? e->row_elems[0] : e->row_elems[1];
  e->dead = (uint8_t*)calloc((size_t)c->n_possible[1], 1);
  e->sat_until = (int32_t*)calloc((size_t)c->n_possible[0], sizeof(int32_t));
  for (int s = 0; s < 2; ++s) {
    e->acc[s] = (double*)calloc((size_t)c->n_possible[s], sizeof(double));
    e->lin_parent[s] = (int32_t*)calloc((size_t)c->n_possible[s], sizeof(int32_t));
    e->lin_live[s] = (int32_t*)calloc((size_t)c->n_possible[s], sizeof(int32_t));
    e->lin_prev[s] = (int32_t*)calloc((size_t)c->n_possible[s], sizeof(int32_t));
    e->lin_alive[s] = (uint8_t*)calloc((size_t)c->n_possible[s], 1);
  }
  e->n_found[0] = c->n_initial[0]; e->n_found[1] = c->n_initial[1];
  e->row_key = (int32_t*)malloc(sizeof(int32_t) * (size_t)(c->n_possible[0] + c->n_possible[1]));
}

void eco_env_free(env_t* e) {
  for (int s = 0; s < 2; ++s) {
    free(e->age[s]); free(e->speed[s]); free(e->termd[s]); free(e->acc[s]);
    free(e->lin_parent[s]); free(e->lin_live[s]); free(e->lin_prev[s]); free(e->lin_alive[s]);
  }
  free(e->gridf); free(e->dead); free(e->row_key); free(e->sat_until);
}

void eco_read_grid(env_t* e, double* out) {
  size_t n = (size_t)e->c->num_obs_channels * e->G * e->G;
  for (size_t i = 0; i < n; ++i) out[i] = (double)e->gridf[i];
}

/* _get_observation + _obs_clip (ECO:700-730): float32, grid channels then the speed plane */
static void eco_get_observation(env_t* e, int s, int id, double* out) {
  const ppg_config* c = e->c;
  const int R = c->obs_range[s], G = e->G, CG = c->num_obs_channels;
  const int off = (R - 1) / 2;
  const int xp = e->x[s][id], yp = e->y[s][id];
  const int xld = xp - off, xhd = xp + off, yld = yp - off, yhd = yp + off;
  const int xlo = clipi(xld, 0, G - 1), xhi = clipi(xhd, 0, G - 1);
  const int ylo = clipi(yld, 0, G - 1), yhi = clipi(yhd, 0, G - 1);
  const int xolo = abs(clipi(xld, -off, 0)), yolo = abs(clipi(yld, -off, 0));
  const int xohi = xolo + (xhi - xlo), yohi = yolo + (yhi - ylo);
  memset(out, 0, sizeof(double) * (size_t)e->row_elems[s]);
  for (int ch = 0; ch < CG; ++ch)
    for (int i = xolo; i <= xohi; ++i)
      for (int j = yolo; j <= yohi; ++j) out[(ch * R + i) * R + j] = (double)*GF(e, ch, xlo + (i - xolo), ylo + (j - yolo));
  if (c->include_speed_in_obs && e->speed[s][id] >= 0.0) { /* ECO:707-711 */
    double norm = (e->speed[s][id] - c->speed_bounds[0]) / (c->speed_bounds[1] - c->speed_bounds[0]);
    if (IS_CAD(c)) norm = e->speed[s][id]; /* CAD:746: the genome value itself */
    const float v = (float)norm; /* assignment into a float32 array */
    for (int i = 0; i < R * R; ++i) out[CG * R * R + i] = (double)v;
  }
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); } /* np.clip */

/* ---- lineage tracking (ECO:1422-1470); parents are of the child's own species ---- */
#define LINEAGE_ON(c) ((c)->trait_mode == PPG_TRAIT_SPEED && ((c)->lineage_reward_coeff[0] != 0.0 || (c)->lineage_reward_coeff[1] != 0.0))
/* _set_lineage_alive_flag + _propagate_lineage_delta: a change of the agent's own alive flag moves the live-descendant
 * count of every ancestor, dead or alive (ECO:1440-1459) */
static void lineage_set_alive(env_t* e, int s, int id, int alive) {
  if (!LINEAGE_ON(e->c) || e->lin_alive[s][id] == (uint8_t)alive) return;
  e->lin_alive[s][id] = (uint8_t)alive;
  for (int a = e->lin_parent[s][id]; a >= 0; a = e->lin_parent[s][a]) e->lin_live[s][a] += alive ? 1 : -1;
}
/* _handle_lineage_birth (ECO:1461-1465) */
static void lineage_birth(env_t* e, int s, int id, int parent) {
  if (!LINEAGE_ON(e->c)) return;
  e->lin_parent[s][id] = parent; e->lin_live[s][id] = 0; e->lin_prev[s][id] = 0; e->lin_alive[s][id] = 0;
  lineage_set_alive(e, s, id, 1);
}

static uint32_t genv(const env_t* e) { return (uint32_t)(e->env_index + e->c->env_index_base); }

/* tape-or-Philox real draws (SURVEY §8c "Tape contents") */
static int take_real(env_t* e, double* out) {
  if (e->tape_reals) {
    if (e->real_pos < e->real_end) { *out = e->tape_reals[e->real_pos++]; return 1; }
    e->status |= PPG_STATUS_TAPE_EXHAUSTED;
  }
  return 0;
}

static void clear_row(env_t* e, int i) {
  e->rew[i] = 0.0; e->has_rew[i] = 0; e->term[i] = -1; e->trunc[i] = -1; e->has_obs[i] = 0;
  e->ate[i] = 0; e->repro[i] = 0; e->newborn[i] = 0; e->carcass[i] = 0; e->born_obs[i] = 0; e->frozen[i] = 0;
}

/* _genome_speed_to_move_rate / _get_agent_move_rate (CAD:556-575): 1/max_cooldown .. 1, linear in the clamped speed */
static double cad_move_rate(const env_t* e, int s, int id) {
  const double sp = e->speed[s][id];
  if (sp < 0.0) return 1.0; /* no genome */
  const double normalized = sp < 0.0 ? 0.0 : (sp > 1.0 ? 1.0 : sp); /* max(0.0, min(1.0, speed)) */
  const double min_rate = 1.0 / (double)e->c->max_cooldown;
  return min_rate + normalized * (1.0 - min_rate);
}

/* reset() (ECO:274-293, 123-223, 1733-1792) from explicit cells and founder speeds */
void eco_env_reset_explicit(env_t* e, const int32_t* cells, const double* founder_speed) {
  const ppg_config* c = e->c;
  const int G = e->G;
  e->current_step = 0;
  memset(e->gridf, 0, sizeof(float) * (size_t)c->num_obs_channels * G * G); /* ECO:222-224 */
  for (int s = 0; s < 2; ++s) {
    memset(e->present[s], 0, (size_t)c->n_possible[s]);
    memset(e->termd[s], 0, (size_t)c->n_possible[s]);
  }
  memset(e->dead, 0, (size_t)c->n_possible[1]); /* ECO:193 */
  e->n_agents = 0;
  if (!IS_MIC(c)) { e->n_found[0] = c->n_initial[0]; e->n_found[1] = c->n_initial[1]; }
  const int nf[2] = {e->n_found[0], e->n_found[1]}; /* trait variants: drawn per episode (MR:189-192) */
  memset(e->sat_until, 0, sizeof(int32_t) * (size_t)c->n_possible[0]); /* MR:1141 */
  int k = 0;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < nf[s]; ++i, ++k) { /* ECO:214-221 + _register_new_agent (ECO:1472-1486) */
      e->agents[e->n_agents++] = KEY(s, i);
      /* _get_initial_age (ECO:1060-1068): founder predators start at the carcass-only threshold */
      e->age[s][i] = (s == 0 && c->carcass_only_predator_age >= 0) ? c->carcass_only_predator_age : 0;
      e->speed[s][i] = c->genome_enabled ? founder_speed[k] : -1.0;
      lineage_birth(e, s, i, -1); /* ECO:1525-1526 */
      /* CAD:1327: random phase of the move accumulator; after the founders' speeds in `founder_speed` */
      if (IS_CAD(c)) e->acc[s][i] = founder_speed[(c->genome_enabled ? nf[0] + nf[1] : 0) + k];
    }
  e->next_idx[0] = nf[0]; /* deque of never-used ids, ascending (ECO:238-258; MR:231-243 lowest unused id) */
  e->next_idx[1] = nf[1];
  k = 0;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < nf[s]; ++i, ++k) { /* ECO:1764-1782 */
      int cx = cells[k] / G, cy = cells[k] % G;
      e->present[s][i] = 1; e->x[s][i] = (int16_t)cx; e->y[s][i] = (int16_t)cy;
      e->energy[s][i] = c->initial_energy[s];
      *GF(e, s, cx, cy) = (float)c->initial_energy[s];
    }
  for (int g = 0; g < c->n_grass; ++g, ++k) { /* ECO:1784-1792 */
    e->gx[g] = (int16_t)(cells[k] / G); e->gy[g] = (int16_t)(cells[k] % G);
    e->ge[g] = c->initial_energy_grass;
    *GF(e, 2, e->gx[g], e->gy[g]) = (float)c->initial_energy_grass;
  }
  e->active[0] = nf[0]; /* ECO:281-282 */
  e->active[1] = nf[1];
  e->cur_num[0] = e->active[0]; e->cur_num[1] = e->active[1];
  eco_ensure_rows(e, e->n_agents);
  for (int i = 0; i < e->n_agents; ++i) { /* ECO:292 */
    clear_row(e, i);
    e->row_key[i] = e->agents[i];
    eco_get_observation(e, KEY_S(e->agents[i]), KEY_ID(e->agents[i]), e->obs + (size_t)i * e->max_row_elems);
    e->has_obs[i] = 1; e->has_rew[i] = 1; e->term[i] = 0; e->trunc[i] = 0;
    if (IS_CAD(c)) e->frozen[i] = !(e->acc[KEY_S(e->agents[i])][KEY_ID(e->agents[i])] + cad_move_rate(e, KEY_S(e->agents[i]), KEY_ID(e->agents[i])) >= 1.0);
  }
  e->n_rows = e->n_agents;
  e->all_term = e->all_trunc = 0;
  e->env_flags = PPG_ENV_RESET;
  e->needs_reset = 0; e->idle = 0; e->status = 0;
  e->spawn_draws = 0;
  for (int k = 0; k < 4; ++k) e->ep_sums[k] = 0.0;
  for (int k = 0; k < 6; ++k) e->ep_events[k] = 0.0;
  e->ep_spawned[0] = e->ep_spawned[1] = 0;
}

/* lockstep reset: founder speeds then cells, from the tape or the Philox streams */
/* founders of the next episode (MR:189-192): two `rng.integers(min, max + 1)` draws, predators first — the first two
 * entries of the episode's cell tape, or the FOUNDERS Philox stream keyed by the episode about to start */
void eco_founder_counts(env_t* e, int from_tape) {
  const ppg_config* c = e->c;
  if (!IS_MIC(c)) { e->n_found[0] = c->n_initial[0]; e->n_found[1] = c->n_initial[1]; return; }
  for (int s = 0; s < 2; ++s) {
    const int lo = c->n_initial_min[s], hi = c->n_initial[s];
    int v;
    if (from_tape && e->tape_cells && e->tape_pos < e->tape_end) v = e->tape_cells[e->tape_pos++];
    else v = lo + (int)ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode + 1, PPG_STREAM_FOUNDERS, (uint32_t)s), (uint32_t)(hi - lo + 1));
    e->n_found[s] = v < lo ? lo : (v > hi ? hi : v);
  }
}

void eco_env_reset_auto(env_t* e) {
  const ppg_config* c = e->c;
  eco_founder_counts(e, 1);
  const int n_f = e->n_found[0] + e->n_found[1], n_total = n_f + c->n_grass;
  const int ncell = e->G * e->G;
  int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_total);
  double* sp = (double*)malloc(sizeof(double) * (size_t)(2 * n_f + 1));
  e->episode += 1;
  e->trait_draws = 0;
  uint8_t sticky = 0;
  if (c->genome_enabled) { /* founder_genome (genome.py:42-46): draws precede the placement (ECO:216-221 before :1752) */
    int k = 0;
    if (e->tape_reals && e->real_pos + n_f <= e->real_end) {
      for (; k < n_f; ++k) sp[k] = clipd(e->tape_reals[e->real_pos++], c->speed_bounds[0], c->speed_bounds[1]);
    } else {
      if (e->tape_reals) sticky |= PPG_STATUS_TAPE_EXHAUSTED;
      for (int s = 0; s < 2; ++s)
        for (int i = 0; i < e->n_found[s]; ++i, ++k) {
          double v = c->founder_speed_mean[s];
          if (c->founder_speed_std[s] > 0) /* genome.py:30-33 */
            v = c->founder_speed_mean[s] + c->founder_speed_std[s] * ppg_draw_normal(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
          sp[k] = clipd(v, c->speed_bounds[0], c->speed_bounds[1]);
        }
    }
  }
  if (IS_CAD(c)) { /* CAD:1327 `rng.uniform(0.0, 1.0)` per founder: from the tape after the speeds, else the trait stream */
    double* acc = sp + (c->genome_enabled ? n_f : 0);
    if (e->tape_reals && e->real_pos + n_f <= e->real_end) {
      for (int k = 0; k < n_f; ++k) acc[k] = e->tape_reals[e->real_pos++];
    } else {
      if (e->tape_reals) sticky |= PPG_STATUS_TAPE_EXHAUSTED;
      for (int k = 0; k < n_f; ++k) acc[k] = ppg_draw_u01(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
    }
  }
  if (e->tape_cells && e->tape_pos + n_total <= e->tape_end) {
    memcpy(cells, e->tape_cells + e->tape_pos, sizeof(int32_t) * (size_t)n_total);
    e->tape_pos += n_total;
  } else {
    if (e->tape_cells) sticky |= PPG_STATUS_TAPE_EXHAUSTED;
    uint8_t* taken = (uint8_t*)calloc((size_t)ncell, 1);
    int n = 0;
    for (uint32_t idx = 0; n < n_total; ++idx) { /* same law as rng.choice(replace=False) (ECO:1752) */
      uint32_t cell = ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_PLACEMENT, idx), (uint32_t)ncell);
      if (!taken[cell]) { taken[cell] = 1; cells[n++] = (int32_t)cell; }
    }
    free(taken);
  }
  eco_env_reset_explicit(e, cells, sp);
  e->status |= sticky;
  free(cells); free(sp);
}

/* speed ** exponent (ECO:559-563): CPython's float power is libm pow(), and so is the oracle's — always.  (The device
 * repeats glibc's pow bit for bit, include/ppg_pow.h; the checker is never bent toward the kernel.) */
static double speed_cost_factor(const env_t* e, double speed) {
  if (IS_MIC(e->c)) return 1.0; /* MR:531-539: cost_per_cell * distance */
  if (speed < 0.0) return 1.0; /* no genome */
  return pow(speed, e->c->move_speed_cost_exponent);
}

static int sgn(int v) { return (v > 0) - (v < 0); }

/* _get_move (ECO:664-695) */
static void eco_get_move(env_t* e, int s, int id, int action, int* nx, int* ny) {
  const ppg_config* c = e->c;
  const int R = c->action_range, d = (R - 1) / 2;
  int dx = action / R - d, dy = action % R - d; /* ECO:225-232 */
  const double sp = e->speed[s][id];
  const int maxd = (sp >= 0.0 && sp >= c->speed_distance_threshold) ? c->fast_max_move_distance : c->slow_max_move_distance; /* ECO:551-557 */
  const int md = abs(dx) > abs(dy) ? abs(dx) : abs(dy);
  if (md > maxd) { dx = sgn(dx) * maxd; dy = sgn(dy) * maxd; } /* ECO:673-677 */
  int x = clipi(e->x[s][id] + dx, 0, e->G - 1), y = clipi(e->y[s][id] + dy, 0, e->G - 1);
  if (*GF(e, s, x, y) > 0) { x = e->x[s][id]; y = e->y[s][id]; } /* ECO:690-692 */
  *nx = x; *ny = y;
}

static int occupied_by_agent(env_t* e, int x, int y) { /* `pos in set(self.agent_positions.values())` (ECO:1136) */
  for (int i = 0; i < e->n_agents; ++i) {
    int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (e->present[s][id] && e->x[s][id] == x && e->y[s][id] == y) return 1;
  }
  return 0;
}

/* _find_available_spawn_position (ECO:732-764); 0 = None */
static int eco_find_spawn(env_t* e, int px, int py, int* ox, int* oy) {
  static const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
  const int G = e->G;
  int vx[4], vy[4], nv = 0;
  for (int k = 0; k < 4; ++k) {
    int x = px + dx[k], y = py + dy[k];
    if (x < 0 || x >= G || y < 0 || y >= G) continue;
    if (!occupied_by_agent(e, x, y)) {
      if (!IS_CAD(e->c)) { *ox = x; *oy = y; return 1; }
      vx[nv] = x; vy[nv] = y; ++nv;
    }
  }
  if (nv > 0) { /* CAD:795 `valid_positions[rng.integers(len(valid_positions))]`: the recorded cell, or a Philox index */
    if (e->tape_cells && e->tape_pos < e->tape_end) {
      int cell = e->tape_cells[e->tape_pos++];
      *ox = cell / G; *oy = cell % G;
      return 1;
    }
    if (e->tape_cells) e->status |= PPG_STATUS_TAPE_EXHAUSTED;
    const uint32_t k = ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_SPAWN, e->spawn_draws++), (uint32_t)nv);
    *ox = vx[k]; *oy = vy[k];
    return 1;
  }
  e->stats[PPG_STAT_SPAWN_FALLBACK]++;
  if (e->tape_cells && e->tape_pos < e->tape_end) {
    int cell = e->tape_cells[e->tape_pos++];
    *ox = cell / G; *oy = cell % G;
    return 1;
  }
  if (e->tape_cells) e->status |= PPG_STATUS_TAPE_EXHAUSTED;
  int n_free = 0;
  for (int cell = 0; cell < G * G; ++cell) n_free += !occupied_by_agent(e, cell / G, cell % G);
  if (n_free == 0) return 0;
  uint32_t k = ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_SPAWN, e->spawn_draws++), (uint32_t)n_free);
  for (int cell = 0; cell < G * G; ++cell)
    if (!occupied_by_agent(e, cell / G, cell % G)) {
      if (k == 0) { *ox = cell / G; *oy = cell % G; return 1; }
      --k;
    }
  return 0;
}

static void capture_obs(env_t* e, int s, int id) { /* self.observations[agent] = self._get_observation(agent) */
  const int i = e->list_index[s][id];
  eco_get_observation(e, s, id, e->obs + (size_t)i * e->max_row_elems);
  e->has_obs[i] = 1;
  /* CAD:577-585,746-753 `_agent_will_move_this_step`: the mask looks one increment ahead of the stored accumulator */
  if (IS_CAD(e->c)) e->frozen[i] = !(e->acc[s][id] + cad_move_rate(e, s, id) >= 1.0);
}

/* _terminate_agent_due_to_age (ECO:1060-1090) */
static void terminate_due_to_age(env_t* e, int s, int id) {
  if (e->termd[s][id] || !e->present[s][id]) return;
  const int i = e->list_index[s][id];
  lineage_set_alive(e, s, id, 0); /* ECO:1069 */
  if (s == 1) { e->dead[id] = 0; e->active[1] = e->active[1] - 1 > 0 ? e->active[1] - 1 : 0; }
  else e->active[0] = e->active[0] - 1 > 0 ? e->active[0] - 1 : 0;
  capture_obs(e, s, id);
  if (!e->has_rew[i]) { e->rew[i] = 0.0; e->has_rew[i] = 1; } /* rewards.get(agent, 0.0) */
  e->term[i] = 1; e->trunc[i] = 0; e->termd[s][id] = 1;
  *GF(e, s, e->x[s][id], e->y[s][id]) = 0;
}

/* _handle_energy_starvation (ECO:766-784) */
static void handle_starvation(env_t* e, int s, int id) {
  const int i = e->list_index[s][id];
  if (s == 1) e->dead[id] = 0;
  lineage_set_alive(e, s, id, 0); /* ECO:770 */
  capture_obs(e, s, id);
  e->rew[i] = 0.0; e->has_rew[i] = 1;
  e->term[i] = 1; e->trunc[i] = 0; e->termd[s][id] = 1;
  *GF(e, s, e->x[s][id], e->y[s][id]) = 0;
  e->active[s] -= 1; /* no guard against a second decrement (quirks 4, 13) */
  e->stats[s == 0 ? PPG_STAT_STARVED_PRED : PPG_STAT_STARVED_PREY]++;
}

/* metabolic_rate ** alpha (MR:751,807): CPython float power = libm pow; 1.0 without a genome */
static double gain_factor(const env_t* e, int s, int id) {
  const double rate = e->speed[s][id] >= 0.0 ? e->speed[s][id] : 1.0;
  return pow(rate, e->c->trait_alpha);
}

/* _apply_cooperative_donation (COOP:537-589): a cooperation_rate share of a positive gain goes, in equal parts, to the
 * live same-species agents within Chebyshev distance cooperation_range; returns what the donor keeps */
static double coop_donation(env_t* e, int s, int id, double gain) {
  const ppg_config* c = e->c;
  if (!c->genome_enabled || gain <= 0.0) return gain;
  const double rate = e->speed[s][id] >= 0.0 ? e->speed[s][id] : 0.0;
  if (rate <= 0.0) return gain;
  const int px = e->x[s][id], py = e->y[s][id], r = c->cooperation_range;
  int n = 0;
  for (int q = 0; q < e->next_idx[s]; ++q) { /* predator_positions / prey_positions: insertion order = ascending id */
    if (q == id || !e->present[s][q] || e->termd[s][q]) continue;
    const int dx = abs(e->x[s][q] - px), dy = abs(e->y[s][q] - py);
    if ((dx > dy ? dx : dy) <= r) ++n;
  }
  if (n == 0) return gain;
  const double total = rate * gain, share = total / n;
  e->ep_events[4 + s] += total;                                       /* COOP:585-586 */
  for (int q = 0; q < e->next_idx[s]; ++q) {
    if (q == id || !e->present[s][q] || e->termd[s][q]) continue;
    const int dx = abs(e->x[s][q] - px), dy = abs(e->y[s][q] - py);
    if ((dx > dy ? dx : dy) > r) continue;
    e->energy[s][q] += share;                                         /* COOP:572 */
    *GF(e, s, e->x[s][q], e->y[s][q]) = (float)e->energy[s][q];       /* COOP:573 */
  }
  return gain - total;
}

/* _handle_prey_engagement of the trait variants (MR:790-834, INV, COOP:812-851) */
static void trait_prey_engagement(env_t* e, int id) {
  const ppg_config* c = e->c;
  const int i = e->list_index[1][id];
  if (e->termd[1][id]) return;
  const int px = e->x[1][id], py = e->y[1][id];
  int grass = -1;
  for (int g = 0; g < c->n_grass; ++g)
    if (e->gx[g] == px && e->gy[g] == py) { grass = g; break; }
  if (grass < 0) { e->rew[i] = c->reward_prey_step; e->has_rew[i] = 1; return; }
  e->ate[i] = 1;
  e->rew[i] = c->reward_prey_eat_grass; e->has_rew[i] = 1;
  const double ge = e->ge[grass];
  double gain = ge;
  if (c->trait_mode == PPG_TRAIT_METABOLIC) gain = ge * gain_factor(e, 1, id);       /* MR:807 */
  else if (c->trait_mode == PPG_TRAIT_COOPERATION) gain = coop_donation(e, 1, id, ge); /* COOP:828 */
  e->energy[1][id] += gain;
  *GF(e, 1, px, py) = (float)e->energy[1][id];
  e->ge[grass] = 0.0;
  *GF(e, 2, px, py) = 0.0f;
  e->stats[PPG_STAT_GRASS_EATEN]++;
}

/* _handle_predator_engagement of the trait variants (MR:720-788, INV, COOP:756-810): the prey is always consumed */
static void trait_predator_engagement(env_t* e, int id) {
  const ppg_config* c = e->c;
  const int i = e->list_index[0][id];
  const int px = e->x[0][id], py = e->y[0][id];
  int caught = -1;
  for (int q = 0; q < e->next_idx[1]; ++q)
    if (e->present[1][q] && e->x[1][q] == px && e->y[1][q] == py) { caught = q; break; }
  if (caught >= 0 && c->satiation_cooldown >= 0 && (c->trait_mode == PPG_TRAIT_METABOLIC || c->trait_mode == PPG_TRAIT_INVESTMENT) &&
      e->current_step < e->sat_until[id]) {
    e->ep_events[3] += 1.0; /* satiation_blocked_catches_predator (MR:739) */
    caught = -1; /* still digesting (MR:734-740) */
  }
  if (caught < 0) { e->rew[i] = c->reward_predator_step; e->has_rew[i] = 1; return; }
  const int j = e->list_index[1][caught];
  e->ate[i] = 1;
  e->rew[i] = c->reward_predator_catch_prey; e->has_rew[i] = 1;
  const double pe = e->energy[1][caught];
  double gain;
  if (c->trait_mode == PPG_TRAIT_COOPERATION) gain = coop_donation(e, 0, id, pe); /* COOP:777 */
  else {
    const double cap = c->max_energy_gain_per_prey;
    const double bite = pe < cap ? pe : cap; /* min(prey_energy, cap): the first argument wins ties (MR:747) */
    gain = c->trait_mode == PPG_TRAIT_METABOLIC ? bite * gain_factor(e, 0, id) : bite; /* MR:751, INV:759 */
  }
  e->energy[0][id] += gain;
  *GF(e, 0, px, py) = (float)e->energy[0][id];
  if (c->satiation_cooldown > 0 && c->trait_mode != PPG_TRAIT_COOPERATION) e->sat_until[id] = e->current_step + c->satiation_cooldown; /* MR:756-757 */
  capture_obs(e, 1, caught);
  e->term[j] = 1; e->termd[1][caught] = 1;
  e->rew[j] = c->penalty_prey_caught; e->has_rew[j] = 1;
  e->trunc[j] = 0;
  e->active[1] -= 1;
  *GF(e, 1, e->x[1][caught], e->y[1][caught]) = 0;
  e->stats[PPG_STAT_EATEN_PREY]++;
}

/* _handle_predator_engagement of the cadence variant (CAD:824-905): the nearest prey within Chebyshev distance 1, the
 * first one in agent_positions order on ties (prey enter the dict in ascending id), eaten whole */
static void cad_predator_engagement(env_t* e, int id) {
  const ppg_config* c = e->c;
  const int i = e->list_index[0][id];
  const int px = e->x[0][id], py = e->y[0][id];
  int caught = -1, best = 0;
  for (int q = 0; q < e->next_idx[1]; ++q) {
    if (!e->present[1][q]) continue;
    const int dx = abs(px - e->x[1][q]), dy = abs(py - e->y[1][q]);
    const int dist = dx > dy ? dx : dy;
    if (dist <= 1 && (caught < 0 || dist < best)) { caught = q; best = dist; }
  }
  if (caught < 0) { e->rew[i] = c->reward_predator_step; e->has_rew[i] = 1; return; }
  const int j = e->list_index[1][caught];
  e->ate[i] = 1;
  e->rew[i] = c->reward_predator_catch_prey; e->has_rew[i] = 1;
  e->energy[0][id] += e->energy[1][caught];                       /* CAD:857-861 */
  *GF(e, 0, px, py) = (float)e->energy[0][id];
  capture_obs(e, 1, caught);                                      /* CAD:869 */
  e->term[j] = 1; e->termd[1][caught] = 1;
  e->rew[j] = c->penalty_prey_caught; e->has_rew[j] = 1;
  e->trunc[j] = 0;
  e->active[1] -= 1;
  *GF(e, 1, e->x[1][caught], e->y[1][caught]) = 0;
  e->stats[PPG_STAT_EATEN_PREY]++;
}

/* _handle_prey_engagement (ECO:885-941) */
static void handle_prey_engagement(env_t* e, int id) {
  const ppg_config* c = e->c;
  if (IS_MIC(c)) { trait_prey_engagement(e, id); return; }
  const int i = e->list_index[1][id];
  if (e->termd[1][id]) return;
  if (e->dead[id]) { e->rew[i] = c->reward_prey_step; e->has_rew[i] = 1; return; }
  const int px = e->x[1][id], py = e->y[1][id];
  int grass = -1;
  for (int g = 0; g < c->n_grass; ++g)
    if (e->gx[g] == px && e->gy[g] == py) { grass = g; break; }
  if (grass >= 0) {
    e->ate[i] = 1;
    e->rew[i] = c->reward_prey_eat_grass; e->has_rew[i] = 1;
    const double ge = e->ge[grass], cap = c->max_energy_gain_per_grass;
    const double bite = ge < cap ? ge : cap; /* min(grass_energy, intake_cap): first argument wins ties */
    e->energy[1][id] += bite;
    *GF(e, 1, px, py) = (float)e->energy[1][id];
    const double rem = ge - bite;
    e->ge[grass] = rem > 0.0 ? rem : 0.0;
    *GF(e, 2, px, py) = (float)e->ge[grass];
    e->stats[PPG_STAT_GRASS_EATEN]++;
  } else {
    e->rew[i] = c->reward_prey_step; e->has_rew[i] = 1;
  }
}

/* _handle_predator_engagement (ECO:786-883) */
static void handle_predator_engagement(env_t* e, int id) {
  const ppg_config* c = e->c;
  if (IS_MIC(c)) { trait_predator_engagement(e, id); return; }
  if (IS_CAD(c)) { cad_predator_engagement(e, id); return; }
  const int i = e->list_index[0][id];
  const int px = e->x[0][id], py = e->y[0][id];
  /* first prey in agent_positions order on the cell (ECO:797-799): prey enter the dict in ascending id
     order; entries of prey terminated earlier in this step are still there (removal is Step 5) */
  int caught = -1;
  for (int q = 0; q < e->next_idx[1]; ++q)
    if (e->present[1][q] && e->x[1][q] == px && e->y[1][q] == py) { caught = q; break; }
  if (caught < 0) { e->rew[i] = c->reward_predator_step; e->has_rew[i] = 1; return; }
  const int was_dead = e->dead[caught];
  if (!was_dead && c->carcass_only_predator_age >= 0 && e->age[0][id] < c->carcass_only_predator_age) { /* ECO:802-804,1070-1078 */
    e->rew[i] = c->reward_predator_step; e->has_rew[i] = 1;
    return;
  }
  const int j = e->list_index[1][caught];
  e->ate[i] = 1;
  e->rew[i] = c->reward_predator_catch_prey; e->has_rew[i] = 1;
  const double pe = e->energy[1][caught], cap = c->max_energy_gain_per_prey;
  const double bite = pe < cap ? pe : cap; /* ECO:812-814 */
  e->energy[0][id] += bite;
  *GF(e, 0, px, py) = (float)e->energy[0][id];
  const double rem = pe - bite;
  if (!was_dead) lineage_set_alive(e, 1, caught, 0); /* the first bite is the prey's death for its lineage (ECO:840-841,848-849) */
  if (rem > 0.0) { /* carcass (ECO:826-845) */
    e->energy[1][caught] = rem;
    *GF(e, 1, e->x[1][caught], e->y[1][caught]) = (float)rem;
    e->dead[caught] = 1;
    /* (if the prey aged out this step, Step 5 removes it and this grid value stays behind: the persistent grid of the
     * oracle keeps it, the device carries it as a ghost cell — ppg_eco.cu header) */
  } else { /* fully eaten (ECO:846-866) */
    capture_obs(e, 1, caught);
    e->term[j] = 1; e->termd[1][caught] = 1;
    e->rew[j] = c->penalty_prey_caught; e->has_rew[j] = 1;
    e->trunc[j] = 0;
    e->active[1] -= 1;
    *GF(e, 1, e->x[1][caught], e->y[1][caught]) = 0;
    e->dead[caught] = 1;
    e->stats[PPG_STAT_EATEN_PREY]++;
  }
}

/* _handle_predator_reproduction / _handle_prey_reproduction (ECO:1092-1275) */
static void handle_reproduction(env_t* e, int s, int id) {
  const ppg_config* c = e->c;
  const int i = e->list_index[s][id];
  if (s == 1 && e->dead[id]) return;                             /* ECO:1192-1193 */
  if (!(e->energy[s][id] >= c->creation_threshold[s])) return;    /* ECO:1100,1195 */
  if (s == 0 && c->trait_mode == PPG_TRAIT_METABOLIC && c->repro_max_ratio >= 0.0 &&
      (double)e->active[0] >= c->repro_max_ratio * (double)e->active[1]) {
    e->ep_events[2] += 1.0; /* reproduction_blocked_due_to_density_predator (MR:852) */
    return; /* density-dependent soft cap (MR:843-854) */
  }
  /* ECO: SystemExit (ECO:1104-1111); the trait variants print a warning and skip the birth (MR:856-864) */
  if (e->next_idx[s] >= c->n_possible[s]) {
    e->status |= PPG_STATUS_ID_POOL_EMPTY;
    if (c->trait_mode != PPG_TRAIT_SPEED) e->ep_events[s] += 1.0; /* reproduction_blocked_due_to_capacity_* (MR:857,943) */
    return;
  }
  if (c->cap_live[s] > 0) { /* device slot capacity (not in the reference) */
    int cnt = 0;
    for (int k = 0; k < e->n_rows; ++k) cnt += (KEY_S(e->row_key[k]) == s);
    if (cnt >= c->cap_live[s]) { e->status |= PPG_STATUS_SLOT_OVERFLOW; return; }
  }
  const int child = e->next_idx[s]++; /* _alloc_new_id: smallest never-used index (ECO:260-272) */
  /* _register_new_agent -> mutate_genome (ECO:1472-1486, genome.py:49-59): the draws precede the spawn search */
  double sp = e->speed[s][id];
  if (c->genome_enabled && c->mutation_rate > 0 && c->mutation_std > 0) {
    double u, d;
    if (!take_real(e, &u)) u = ppg_draw_u01(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
    if (u < c->mutation_rate) {
      if (!take_real(e, &d)) d = c->mutation_std * ppg_draw_normal(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
      sp = clipd(sp + d, c->speed_bounds[0], c->speed_bounds[1]);
    }
  }
  double acc0 = 0.0;
  if (IS_CAD(c) && !take_real(e, &acc0)) /* CAD:1327: after the genome, before the spawn search */
    acc0 = ppg_draw_u01(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
  int sx, sy;
  if (!eco_find_spawn(e, e->x[s][id], e->y[s][id], &sx, &sy)) { /* RuntimeError (ECO:1142-1143) */
    e->status |= PPG_STATUS_NO_SPAWN_CELL;
    e->next_idx[s]--; /* the reference dies here; the lockstep layer drops the birth but keeps the consumed draws */
    return;
  }
  const int ci = e->n_rows;
  eco_ensure_rows(e, ci + 2);
  e->agents[e->n_agents++] = KEY(s, child); /* ECO:1113 */
  e->row_key[e->n_rows++] = KEY(s, child);
  clear_row(e, ci);
  e->list_index[s][child] = ci;
  e->newborn[ci] = 1;
  e->age[s][child] = 0;
  e->speed[s][child] = c->genome_enabled ? sp : -1.0;
  e->termd[s][child] = 0;
  if (s == 1) e->dead[child] = 0;
  e->present[s][child] = 1; e->x[s][child] = (int16_t)sx; e->y[s][child] = (int16_t)sy; /* ECO:1145-1146 */
  double child_e = c->initial_energy[s];               /* ECO:1148-1149 */
  if (c->trait_mode == PPG_TRAIT_INVESTMENT) {         /* INV:546-557,890,977: parent energy * the PARENT's fraction */
    const double fraction = e->speed[s][id] >= 0.0 ? e->speed[s][id] : c->founder_speed_mean[s];
    child_e = e->energy[s][id] * fraction;
  }
  if (s == 0) e->sat_until[child] = 0;                 /* MR:1141 */
  e->acc[s][child] = acc0;
  lineage_birth(e, s, child, id);                      /* ECO:1525-1526 (in _register_new_agent) */
  e->energy[s][child] = child_e;
  e->energy[s][id] -= child_e;                         /* ECO:1150 */
  *GF(e, s, sx, sy) = (float)child_e;                  /* ECO:1154 */
  *GF(e, s, e->x[s][id], e->y[s][id]) = (float)e->energy[s][id]; /* ECO:1155 */
  e->active[s] += 1;                                   /* ECO:1157 */
  e->ep_spawned[s] += 1;                               /* ECO:1168 offspring_count of the parent's record */
  e->rew[ci] = 0.0; e->has_rew[ci] = 1;                /* ECO:1160 */
  e->rew[i] = c->reproduction_reward[s]; e->has_rew[i] = 1; /* ECO:1161 */
  e->repro[i] = 1;                                     /* ECO:1168 agent_offspring_counts[agent] += 1 */
  capture_obs(e, s, child);                            /* ECO:1179 */
  e->born_obs[ci] = 1;
  e->term[ci] = 0; e->trunc[ci] = 0;
  e->stats[s == 0 ? PPG_STAT_BIRTHS_PRED : PPG_STAT_BIRTHS_PREY]++;
}

/*
 * step(action_dict) (ECO:295-507).  The action dict is given in the caller's iteration order;
 * only the movement loop iterates it (ECO:632).  Returns -1 if a key is not a live agent.
 */
int eco_env_step(env_t* e, int n_act, const int32_t* a_s, const int32_t* a_id, const int32_t* a_val) {
  const ppg_config* c = e->c;
  e->env_flags = 0;
  const int n0 = e->n_agents;
  eco_ensure_rows(e, n0 + 1);
  e->n_rows = n0;
  for (int i = 0; i < n0; ++i) {
    clear_row(e, i);
    e->row_key[i] = e->agents[i];
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    e->list_index[s][id] = i;
    e->termd[s][id] = 0; /* self.terminations = {} (ECO:297) */
  }
  for (int k = 0; k < n_act; ++k)
    if (a_id[k] < 0 || a_id[k] >= c->n_possible[a_s[k]] || !e->present[a_s[k]][a_id[k]]) return -1;
  e->stats[PPG_STAT_ENV_STEPS]++;
  e->stats[PPG_STAT_AGENT_STEPS] += n0;

  /* Step 1: _apply_time_step_update (ECO:582-616), over list(self.agents) */
  for (int i = 0; i < n0; ++i) {
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    double decay = c->energy_loss[s];
    if (c->trait_mode == PPG_TRAIT_METABOLIC) decay = c->energy_loss[s] * (e->speed[s][id] >= 0.0 ? e->speed[s][id] : 1.0); /* MR:555-561 */
    if (IS_CAD(c) && c->genome_enabled && c->metabolic_speed_coeff > 0.0 && e->speed[s][id] >= 0.0)
      decay *= 1.0 + c->metabolic_speed_coeff * e->speed[s][id]; /* CAD:626-633 */
    e->energy[s][id] -= decay;
    *GF(e, s, e->x[s][id], e->y[s][id]) = (float)e->energy[s][id];
    if (!(s == 1 && e->dead[id])) e->age[s][id] += 1; /* carcasses do not age (ECO:600-601) */
  }
  for (int i = 0; i < n0; ++i) { /* aged_out_agents, in self.agents order (ECO:602-603,615-616) */
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (s == 1 && e->dead[id]) continue;
    if (c->max_agent_age[s] >= 0 && e->age[s][id] >= c->max_agent_age[s]) terminate_due_to_age(e, s, id); /* ECO:1052-1058 */
  }
  /* Step 2: _regenerate_grass_energy (ECO:618-626) */
  for (int g = 0; g < c->n_grass; ++g) {
    const double v = e->ge[g] + c->energy_gain_grass;
    e->ge[g] = v < c->max_energy_grass ? v : c->max_energy_grass;
    *GF(e, 2, e->gx[g], e->gy[g]) = (float)e->ge[g];
  }
  /* Step 3: _process_agent_movements (ECO:628-662), action-dict order */
  for (int k = 0; k < n_act; ++k) {
    const int s = a_s[k], id = a_id[k];
    int act = a_val[k];
    if (!e->present[s][id] || e->termd[s][id]) continue; /* ECO:633-634 */
    if (s == 1 && e->dead[id]) continue;                  /* ECO:636-637 */
    if (IS_CAD(c)) { /* cadence gate (CAD:674-681): frozen agents keep their place, the accumulator still advances */
      const double a = e->acc[s][id] + cad_move_rate(e, s, id);
      if (a < 1.0) { e->acc[s][id] = a; continue; }
      e->acc[s][id] = a - 1.0;
    }
    if (act < 0 || act >= c->action_range * c->action_range) { e->status |= PPG_STATUS_BAD_ACTION; act = (c->action_range * c->action_range) / 2; }
    const int ox = e->x[s][id], oy = e->y[s][id];
    int nx, ny;
    eco_get_move(e, s, id, act, &nx, &ny);
    /* _get_movement_energy_cost (ECO:565-573): np.linalg.norm of an integer vector = sqrt(dx^2 + dy^2) in float64 */
    const double ddx = (double)(nx - ox), ddy = (double)(ny - oy);
    const double dist = sqrt(ddx * ddx + ddy * ddy);
    double cost = 0.0;
    if (dist > 0) cost = c->move_cost_per_cell[s] * dist * speed_cost_factor(e, e->speed[s][id]);
    e->energy[s][id] -= cost;
    e->ep_sums[s] += dist; e->ep_sums[2 + s] += cost; /* ECO:659-660 */
    *GF(e, s, ox, oy) = 0;
    *GF(e, s, nx, ny) = (float)e->energy[s][id];
    e->x[s][id] = (int16_t)nx; e->y[s][id] = (int16_t)ny;
  }
  /* Step 4a: starvation over tuple(agent_energies.items()) = self.agents order (ECO:311-316) */
  for (int i = 0; i < n0; ++i) {
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (e->energy[s][id] <= 0) handle_starvation(e, s, id);
  }
  /* Step 4b: prey engagements (ECO:318-323) */
  for (int i = 0; i < n0; ++i) {
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (s == 1 && !e->termd[1][id]) handle_prey_engagement(e, id);
  }
  /* Step 4c: predator engagements (ECO:325-330) */
  for (int i = 0; i < n0; ++i) {
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (s == 0 && !e->termd[0][id]) handle_predator_engagement(e, id);
  }
  /* Step 5: removals (ECO:332-351) */
  {
    int w = 0;
    for (int i = 0; i < n0; ++i) {
      const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
      if (e->termd[s][id]) e->present[s][id] = 0; else e->agents[w++] = e->agents[i];
    }
    e->n_agents = w;
  }
  /* Step 6: reproduction over snapshots, predators then prey (ECO:353-367) */
  {
    const int n_live = e->n_agents;
    for (int sp = 0; sp < 2; ++sp)
      for (int i = 0; i < n_live; ++i) {
        const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
        if (s == sp && e->energy[s][id] >= c->creation_threshold[s]) handle_reproduction(e, s, id);
      }
  }
  /* Step 6.5: _apply_lineage_survival_rewards (ECO:369-370,943-984): every living agent (carcasses and newborns included)
   * is paid coeff * (change of its live-descendant count since the last step) on top of what it already has */
  if (LINEAGE_ON(c))
    for (int i = 0; i < e->n_agents; ++i) {
      const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
      const int li = e->list_index[s][id];
      if (!e->has_rew[li]) { e->rew[li] = 0.0; e->has_rew[li] = 1; } /* rewards.setdefault(agent_id, 0.0) */
      const int delta = e->lin_live[s][id] - e->lin_prev[s][id];
      e->lin_prev[s][id] = e->lin_live[s][id];
      if (delta == 0) continue;
      const double reward = c->lineage_reward_coeff[s] * (double)delta;
      if (reward != 0) e->rew[li] = e->rew[li] + reward; /* ECO:965-966 */
    }
  /* Step 7: outputs (ECO:372-424) */
  const int episode_done = e->active[1] <= 0 || e->active[0] <= 0; /* ECO:392 */
  for (int i = 0; i < e->n_rows; ++i) {
    const int s = KEY_S(e->row_key[i]), id = KEY_ID(e->row_key[i]);
    if (e->term[i] == 1) continue; /* ended: keeps the observation captured at its end (ECO:415-422) */
    /* live agents */
    if (!e->has_rew[i]) { e->rew[i] = 0.0; e->has_rew[i] = 1; }
    if (episode_done) {
      e->term[i] = 1; e->trunc[i] = 0;                       /* ECO:393-398 */
      if (!e->born_obs[i]) capture_obs(e, s, id);            /* ECO:417-420: newborns keep the capture of ECO:1179 */
    } else {
      e->term[i] = 0; e->trunc[i] = 0;
      capture_obs(e, s, id);                                 /* ECO:424 */
    }
    if (s == 1 && e->dead[id]) e->carcass[i] = 1;
  }
  e->all_term = (uint8_t)episode_done;
  e->all_trunc = 0;
  e->current_step += 1; /* ECO:450 */
  if (e->current_step >= c->max_steps && !episode_done) { /* ECO:452-501 */
    for (int i = 0; i < e->n_rows; ++i) {
      const int s = KEY_S(e->row_key[i]), id = KEY_ID(e->row_key[i]);
      if (e->present[s][id]) { e->trunc[i] = 1; e->term[i] = 0; capture_obs(e, s, id); }
    }
    e->all_trunc = 1;
  }
  if (e->all_term) e->env_flags |= PPG_ENV_TERMINATED;
  if (e->all_trunc) e->env_flags |= PPG_ENV_TRUNCATED;
  e->cur_num[0] = e->active[0]; e->cur_num[1] = e->active[1];
  if (episode_done || e->all_trunc) e->n_agents = 0; /* self.agents = [] (ECO:499,503) */
  return 0;
}
