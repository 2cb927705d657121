/*
 * ppg_oracle_stag.c — CPU oracle of the STAG env (TEST INFRASTRUCTURE ONLY, see ppg_oracle.h).
 *
 * Sequential restatement of
 *   STAG = predpreygrass/evolutionary/stag_hunt_forward_view_nature_nurture/predpreygrass_rllib_env.py
 * with the reference's data structures: dicts keyed by agent id -> per-id arrays (flat ids, include/ppg.h),
 * `self.agents` (whose order is also the insertion order of agent_positions / agent_energies /
 * predator_positions / prey_positions: all append at birth and delete at death), the persistent FLOAT32
 * five-channel grid (STAG:385-387), predator facing and cooperation trait, join intents, agents_just_ate.
 * Every block cites the lines it follows.
 *
 * Not restated (host-side analytics that never feed back into the step): agent_event_log, agent_stats_*,
 * per_step_agent_data, infos, team_capture_events.  Walls / line of sight (STAG:892-925,2196-2217) are rejected
 * by the config layer (the BASELINE config has none).
 *
 * Pinned against golden trajectories recorded from the unmodified reference
 * (tests/golden/make_golden_stag.py -> tests/golden/stag_*.npz, tests/test_oracle_golden_stag.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ppg_philox.h"
#include "ppg_oracle_int.h"

enum { CH_WALL = 0, CH_PRED = 1, CH_PREY1 = 2, CH_PREY2 = 3, CH_GRASS = 4 }; /* STAG:112-117 */

static inline float* GF(env_t* e, int ch, int x, int y) { return &e->gridf[((size_t)ch * e->G + x) * e->G + y]; }
static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int type_of(const ppg_config* c, int s, int id) { return id >= c->n_possible_t[s][0]; } /* 0 = type_1, 1 = type_2 */
static inline int channel_of(const ppg_config* c, int s, int id) { return s == 0 ? CH_PRED : (type_of(c, 1, id) ? CH_PREY2 : CH_PREY1); } /* STAG:2082-2092 */
static uint32_t genv(const env_t* e) { return (uint32_t)(e->env_index + e->c->env_index_base); }

void stag_env_alloc(env_t* e) {
  const ppg_config* c = e->c;
  eco_env_alloc(e); /* float32 grid, ages, per-step termination marks, row keys */
  e->facing = (int8_t*)calloc((size_t)c->n_possible[0] + 1, 1);
  e->trait = (double*)calloc((size_t)c->n_possible[0] + 1, sizeof(double));
  e->join = (uint8_t*)calloc((size_t)c->n_possible[0] + 1, 1);
  e->wallmap = (uint8_t*)calloc((size_t)e->G * e->G, 1); /* _create_wall_positions (STAG:2107-2127): static per config */
  for (int k = 0; k < c->n_walls; ++k) e->wallmap[c->wall_cells[k]] = 1;
}

void stag_env_free(env_t* e) {
  eco_env_free(e);
  free(e->facing); free(e->trait); free(e->join); free(e->wallmap);
}

/* _predator_facing_options (STAG:197-206) */
static const int FDX[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, FDY[8] = {-1, 0, 1, -1, 1, -1, 0, 1};

/* _get_observation + _get_predator_view_center + _obs_clip (STAG:944-1008) */
static void stag_get_observation(env_t* e, int s, int id, double* out) {
  const ppg_config* c = e->c;
  const int R = c->obs_range[s], G = e->G, CG = c->num_obs_channels;
  const int off = (R - 1) / 2;
  int xp = e->x[s][id], yp = e->y[s][id];
  if (s == 0) { /* forward view: window centre shifted by facing * offset (STAG:944-951) */
    const int f = e->facing[id];
    if (f >= 0) { xp += FDX[f] * off; yp += FDY[f] * off; }
  }
  const int xld = xp - off, xhd = xp + off, yld = yp - off, yhd = yp + off;
  const int xlo = clipi(xld, 0, G - 1), xhi = clipi(xhd, 0, G - 1);
  const int ylo = clipi(yld, 0, G - 1), yhi = clipi(yhd, 0, G - 1);
  const int xolo = abs(clipi(xld, -off, 0)), yolo = abs(clipi(yld, -off, 0)); /* saturates */
  const int xohi = xolo + (xhi - xlo), yohi = yolo + (yhi - ylo);
  memset(out, 0, sizeof(double) * (size_t)e->row_elems[s]);
  for (int ch = 0; ch < CG; ++ch)
    for (int i = xolo; i <= xohi; ++i)
      for (int j = yolo; j <= yohi; ++j) out[(ch * R + i) * R + j] = (double)*GF(e, ch, xlo + (i - xolo), ylo + (j - yolo));
  /* the line-of-sight masks are computed once in __init__, when wall_positions is still empty (STAG:175,408-412): all
   * ones, so masking the dynamic channels changes nothing (STAG:989-992) and the visibility channel is a plane of ones
   * over the WHOLE window, clipped or not (STAG:993-994) */
  if (c->include_visibility_channel)
    for (int i = 0; i < R * R; ++i) out[CG * R * R + i] = 1.0;
}

static int is_wall(const env_t* e, int x, int y) { return e->wallmap[x * e->G + y]; }

/* _line_of_sight_clear (STAG:892-925): integer Bresenham walk from start to end, no wall strictly between */
static int los_clear(const env_t* e, int x0, int y0, int x1, int y1) {
  const int dx = abs(x1 - x0), dy = abs(y1 - y0);
  int x = x0, y = y0;
  const int sx = x1 > x0 ? 1 : -1, sy = y1 > y0 ? 1 : -1;
  if (dx >= dy) {
    double err = dx / 2.0;
    while (x != x1) {
      if (!(x == x0 && y == y0) && !(x == x1 && y == y1) && is_wall(e, x, y)) return 0;
      err -= dy;
      if (err < 0) { y += sy; err += dx; }
      x += sx;
    }
  } else {
    double err = dy / 2.0;
    while (y != y1) {
      if (!(x == x0 && y == y0) && !(x == x1 && y == y1) && is_wall(e, x, y)) return 0;
      err -= dx;
      if (err < 0) { x += sx; err += dy; }
      y += sy;
    }
  }
  return 1;
}

static int take_real(env_t* e, double* out) {
  if (e->tape_reals) {
    if (e->real_pos < e->real_end) { *out = e->tape_reals[e->real_pos++]; return 1; }
    e->status |= PPG_STATUS_TAPE_EXHAUSTED;
  }
  return 0;
}
static int take_int(env_t* e, int* out) {
  if (e->tape_cells) {
    if (e->tape_pos < e->tape_end) { *out = e->tape_cells[e->tape_pos++]; return 1; }
    e->status |= PPG_STATUS_TAPE_EXHAUSTED;
  }
  return 0;
}

static void clear_row(env_t* e, int i) {
  e->rew[i] = 0.0; e->has_rew[i] = 0; e->term[i] = -1; e->trunc[i] = -1; e->has_obs[i] = 0;
  e->ate[i] = 0; e->repro[i] = 0; e->newborn[i] = 0; e->carcass[i] = 0; e->born_obs[i] = 0;
}

static double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); } /* _clip_trait (STAG:1080-1082): min(max(v, 0), 1) */

/* reset() (STAG:414-430, 268-412, 2095-2193) from explicit cells, founder facings and raw founder traits */
void stag_env_reset_explicit(env_t* e, const int32_t* cells, const int32_t* facing, const double* trait_raw) {
  const ppg_config* c = e->c;
  const int G = e->G;
  e->current_step = 0;
  memset(e->gridf, 0, sizeof(float) * (size_t)c->num_obs_channels * G * G);
  for (int s = 0; s < 2; ++s) {
    memset(e->present[s], 0, (size_t)c->n_possible[s]);
    memset(e->termd[s], 0, (size_t)c->n_possible[s]);
  }
  memset(e->capture, 0, sizeof e->capture);
  e->capture_real[0] = e->capture_real[1] = e->capture_real[2] = 0.0;
  e->n_agents = 0;
  for (int s = 0; s < 2; ++s)      /* STAG:373-380: predators type 1, type 2, then prey type 1, type 2 */
    for (int t = 0; t < 2; ++t) {
      for (int i = 0; i < c->n_initial_t[s][t]; ++i) e->agents[e->n_agents++] = KEY(s, (t ? c->n_possible_t[s][0] : 0) + i);
      e->next_idx_t[s][t] = c->n_initial_t[s][t]; /* pools hold the never-used ids in ascending order (STAG:2259-2291) */
    }
  for (int cell = 0; cell < G * G; ++cell) /* _place_walls (STAG:2152-2159) */
    if (e->wallmap[cell]) *GF(e, 0, cell / G, cell % G) = 1.0f;
  int k = 0, kp = 0;
  for (int i = 0; i < e->n_agents; ++i, ++k) { /* _place_predators / _place_prey (STAG:2162-2183) */
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    const int cx = cells[k] / G, cy = cells[k] % G;
    e->present[s][id] = 1; e->x[s][id] = (int16_t)cx; e->y[s][id] = (int16_t)cy;
    e->age[s][id] = 0; /* STAG:2255-2257 */
    if (s == 0) {
      e->facing[id] = (int8_t)facing[kp];
      e->trait[id] = c->coop_trait_enabled ? clip01(trait_raw[kp]) : 1.0; /* STAG:1084-1088 */
      ++kp;
      e->energy[0][id] = c->initial_energy[0];
    } else {
      e->energy[1][id] = c->initial_energy_prey_t[type_of(c, 1, id)];
    }
    *GF(e, channel_of(c, s, id), cx, cy) = (float)e->energy[s][id];
  }
  for (int g = 0; g < c->n_grass; ++g, ++k) { /* _place_grass (STAG:2186-2193) */
    e->gx[g] = (int16_t)(cells[k] / G); e->gy[g] = (int16_t)(cells[k] % G);
    e->ge[g] = c->initial_energy_grass;
    *GF(e, CH_GRASS, e->gx[g], e->gy[g]) = (float)c->initial_energy_grass;
  }
  e->active[0] = c->n_initial[0]; e->active[1] = c->n_initial[1]; /* STAG:423-424 */
  e->cur_num[0] = e->active[0]; e->cur_num[1] = e->active[1];
  eco_ensure_rows(e, e->n_agents);
  for (int i = 0; i < e->n_agents; ++i) { /* STAG:429 */
    clear_row(e, i);
    e->row_key[i] = e->agents[i];
    stag_get_observation(e, KEY_S(e->agents[i]), KEY_ID(e->agents[i]), e->obs + (size_t)i * e->max_row_elems);
    e->has_obs[i] = 1; e->has_rew[i] = 1; e->term[i] = 0; e->trunc[i] = 0;
  }
  e->n_rows = e->n_agents;
  e->all_term = e->all_trunc = 0;
  e->env_flags = PPG_ENV_RESET;
  e->needs_reset = 0; e->idle = 0; e->status = 0;
  e->spawn_draws = 0; e->capture_draws = 0;
  e->facing_draws = (uint32_t)c->n_initial[0]; /* the founders own facing draws 0..n-1 of the episode */
}

/* lockstep reset: cells, facings, traits from the tape or the Philox streams */
void stag_env_reset_auto(env_t* e) {
  const ppg_config* c = e->c;
  const int n_f = c->n_initial[0] + c->n_initial[1], n_total = n_f + c->n_grass, n_pred = c->n_initial[0];
  const int ncell = e->G * e->G;
  int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_total + n_pred + 1));
  int32_t* facing = cells + n_total;
  double* tr = (double*)malloc(sizeof(double) * (size_t)(n_pred + 1));
  e->episode += 1;
  e->trait_draws = 0;
  uint8_t sticky = 0;
  if (e->tape_cells && e->tape_pos + n_total + n_pred <= e->tape_end) {
    memcpy(cells, e->tape_cells + e->tape_pos, sizeof(int32_t) * (size_t)(n_total + n_pred));
    e->tape_pos += n_total + n_pred;
  } else {
    if (e->tape_cells) sticky |= PPG_STATUS_TAPE_EXHAUSTED;
    uint8_t* taken = (uint8_t*)calloc((size_t)ncell, 1);
    int n = 0;
    for (uint32_t idx = 0; n < n_total; ++idx) { /* same law as rng.choice(replace=False) (STAG:2140) */
      uint32_t cell = ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_PLACEMENT, idx), (uint32_t)ncell);
      if (!taken[cell] && !e->wallmap[cell]) { taken[cell] = 1; cells[n++] = (int32_t)cell; } /* free_non_wall_indices (STAG:2138) */
    }
    free(taken);
    for (int k = 0; k < n_pred; ++k) /* _random_predator_facing (STAG:939-942) */
      facing[k] = (int32_t)ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_FACING, (uint32_t)k), 8u);
  }
  if (c->coop_trait_enabled) { /* _sample_initial_predator_trait (STAG:1084-1088) */
    if (e->tape_reals && e->real_pos + n_pred <= e->real_end) {
      for (int k = 0; k < n_pred; ++k) tr[k] = e->tape_reals[e->real_pos++];
    } else {
      if (e->tape_reals) sticky |= PPG_STATUS_TAPE_EXHAUSTED;
      for (int k = 0; k < n_pred; ++k) {
        double v = c->coop_trait_init_mean;
        if (c->coop_trait_init_std > 0)
          v = c->coop_trait_init_mean + c->coop_trait_init_std * ppg_draw_normal(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
        tr[k] = v;
      }
    }
  }
  stag_env_reset_explicit(e, cells, facing, tr);
  e->status |= sticky;
  free(cells); free(tr);
}

static void remove_from_agents(env_t* e, int32_t key) { /* self.agents.remove(agent_id) (STAG:2233-2237) */
  int w = 0;
  for (int i = 0; i < e->n_agents; ++i)
    if (e->agents[i] != key) e->agents[w++] = e->agents[i];
  e->n_agents = w;
}

static double death_penalty(const ppg_config* c, int s, int id) { /* STAG:1999-2006 */
  return s == 0 ? c->death_penalty[0] : c->death_penalty[1 + type_of(c, 1, id)];
}

/* _handle_energy_starvation (STAG:1046-1067); the observation captured here is overwritten at STAG:551 */
static void handle_starvation(env_t* e, int s, int id) {
  const ppg_config* c = e->c;
  const int i = e->list_index[s][id];
  if (!e->has_rew[i]) { e->rew[i] = 0.0; e->has_rew[i] = 1; }
  const double pen = death_penalty(c, s, id);
  if (pen != 0.0) e->rew[i] += pen;
  e->term[i] = 1; e->trunc[i] = 0; e->termd[s][id] = 1;
  *GF(e, channel_of(c, s, id), e->x[s][id], e->y[s][id]) = 0;
  e->active[s] -= 1;
  e->present[s][id] = 0; /* _remove_agent_from_state (STAG:2219-2237) */
  remove_from_agents(e, KEY(s, id));
  e->stats[s == 0 ? PPG_STAT_STARVED_PRED : PPG_STAT_STARVED_PREY]++;
}

/* Python's builtin sum() over floats: CPython >= 3.12 (the interpreter the golden files were recorded with) adds the
 * first item exactly (0 + x) and the rest with Neumaier's compensated summation (bltinmodule.c, builtin_sum_impl) */
static double py_sum(const double* v, int n) {
  if (n == 0) return 0.0;
  double f = v[0], c = 0.0;
  for (int k = 1; k < n; ++k) {
    const double x = v[k], t = f + x;
    if (fabs(f) >= fabs(x)) c += (f - t) + x; else c += (x - t) + f;
    f = t;
  }
  if (c != 0.0 && isfinite(c)) f += c;
  return f;
}

/* _compute_team_capture_success (STAG:1117-1150) */
static int capture_success(env_t* e, const int* joiners, int nj, double prey_energy, double* prob_out, double* ratio_out) {
  const ppg_config* c = e->c;
  const double m = prey_energy + c->team_capture_margin;
  const double difficulty = m > 1e-8 ? m : 1e-8; /* max(a, b): a unless b > a */
  if (!c->coop_trait_enabled) {
    double* v = (double*)malloc(sizeof(double) * (size_t)nj);
    for (int k = 0; k < nj; ++k) v[k] = e->energy[0][joiners[k]];
    const double total = py_sum(v, nj); /* `sum(...)` (STAG:1123) */
    free(v);
    const double ratio = total / difficulty;
    const int ok = ratio > 1.0;
    *prob_out = ok ? 1.0 : 0.0; *ratio_out = ratio;
    return ok;
  }
  const double nw = c->team_capture_nature_weight;
  double total = 0.0;
  for (int k = 0; k < nj; ++k) {
    const double factor = (1.0 - nw) + nw * e->trait[joiners[k]];
    total += e->energy[0][joiners[k]] * factor;
  }
  const double ratio = total / difficulty;
  const double ex = ratio > 0.0 ? ratio : 0.0;
  const double base = 1.0 - c->team_capture_base_success_p0;
  const double pw = pow(base, ex); /* CPython `**` = libm pow (STAG:1137); the device repeats glibc's pow, include/ppg_pow.h */
  const double base_prob = 1.0 - pw;
  double prob = base_prob > c->team_capture_min_success_prob ? base_prob : c->team_capture_min_success_prob;
  if (prob > 1.0) prob = 1.0;
  const int force = ratio >= c->team_capture_force_success_ratio;
  int ok;
  if (c->team_capture_success_model == PPG_CAPTURE_DETERMINISTIC) {
    ok = ratio > 1.0;
    prob = ok ? 1.0 : 0.0;
  } else if (c->team_capture_success_model == PPG_CAPTURE_PROBABILISTIC || !force) {
    double u;
    if (!take_real(e, &u)) u = ppg_draw_u01(e->seed_key, genv(e), e->episode, PPG_STREAM_CAPTURE, &e->capture_draws);
    ok = u < prob;
  } else {
    ok = 1; /* hybrid, force_success: no draw (short-circuit `or`, STAG:1148) */
  }
  *prob_out = prob; *ratio_out = ratio;
  return ok;
}

/* _handle_team_capture (STAG:1152-1442); 1 = prey captured */
static int handle_team_capture(env_t* e, int prey) {
  const ppg_config* c = e->c;
  if (e->termd[1][prey]) return 0;
  const int px = e->x[1][prey], py = e->y[1][prey];
  int* joiners = (int*)malloc(sizeof(int) * (size_t)(e->n_agents + 1));
  int* riders = (int*)malloc(sizeof(int) * (size_t)(e->n_agents + 1));
  int nj = 0, nr = 0;
  for (int i = 0; i < e->n_agents; ++i) { /* predator_positions order; Moore neighbourhood (STAG:1069-1078) */
    if (KEY_S(e->agents[i]) != 0) continue;
    const int id = KEY_ID(e->agents[i]);
    if (e->termd[0][id]) continue;
    const int dx = abs(e->x[0][id] - px), dy = abs(e->y[0][id] - py);
    if ((dx > dy ? dx : dy) > 1) continue;
    if (e->ate[e->list_index[0][id]]) continue;             /* h not in agents_just_ate (STAG:1158) */
    if (e->join[id] != 0) joiners[nj++] = id; else riders[nr++] = id; /* join intent defaults to True (STAG:1163) */
  }
  if (nj == 0) { free(joiners); free(riders); return 0; }   /* STAG:1160-1166 */
  const double prey_energy = e->energy[1][prey];
  const double join_cost = c->team_capture_join_cost;
  const int rabbit = type_of(c, 1, prey);
  double prob, ratio;
  const int success = capture_success(e, joiners, nj, prey_energy, &prob, &ratio);
  e->capture_real[0] = prob; e->capture_real[1] = ratio; e->capture_real[2] += prob; /* STAG:1176-1179 */
  e->capture[8] += 1;
  e->stats[PPG_STAT_CAPTURE_ATTEMPTS]++;
  if (!success) {
    if (join_cost != 0.0)
      for (int k = 0; k < nj; ++k) { /* STAG:1185-1192 */
        const int pid = joiners[k];
        e->energy[0][pid] -= join_cost;
        *GF(e, CH_PRED, e->x[0][pid], e->y[0][pid]) = (float)e->energy[0][pid];
      }
    for (int k = 0; k < nj; ++k) /* STAG:1238-1240 */
      if (e->energy[0][joiners[k]] <= 0 && !e->termd[0][joiners[k]]) handle_starvation(e, 0, joiners[k]);
    e->capture[1] += 1; e->capture[rabbit ? 7 : 5] += 1;
    if (nj > 1) e->capture[3] += 1;
    free(joiners); free(riders);
    return 0;
  }
  e->capture[0] += 1; e->capture[rabbit ? 6 : 4] += 1; /* STAG:1264-1271 */
  if (nj > 1) e->capture[2] += 1;
  e->capture[9] += nj;
  double* snap = (double*)malloc(sizeof(double) * (size_t)nj);
  for (int k = 0; k < nj; ++k) snap[k] = e->energy[0][joiners[k]];
  const double total_helper = py_sum(snap, nj); /* `sum(helper_energy_snapshot.values())` (STAG:1273-1274) */
  const double scav_frac = nr ? c->team_capture_scavenger_fraction : 0.0;
  const double scav_total = prey_energy * scav_frac;
  const double pool = prey_energy - scav_total;
  for (int k = 0; k < nj; ++k) { /* STAG:1279-1297 */
    const int pid = joiners[k], i = e->list_index[0][pid];
    e->ate[i] = 1;
    double share;
    if (c->team_capture_equal_split) share = pool / (double)nj;
    else share = total_helper > 0 ? pool * (snap[k] / total_helper) : 0.0;
    e->energy[0][pid] += share;
    if (join_cost != 0.0) e->energy[0][pid] -= join_cost;
    *GF(e, CH_PRED, e->x[0][pid], e->y[0][pid]) = (float)e->energy[0][pid];
    if (!e->has_rew[i]) { e->rew[i] = 0.0; e->has_rew[i] = 1; }
  }
  const double scav_share = nr ? scav_total / (double)nr : 0.0; /* STAG:1338-1348 */
  if (scav_share != 0.0)
    for (int k = 0; k < nr; ++k) {
      const int pid = riders[k], i = e->list_index[0][pid];
      e->ate[i] = 1;
      e->energy[0][pid] += scav_share;
      *GF(e, CH_PRED, e->x[0][pid], e->y[0][pid]) = (float)e->energy[0][pid];
      if (!e->has_rew[i]) { e->rew[i] = 0.0; e->has_rew[i] = 1; }
    }
  if (join_cost != 0.0) /* STAG:1402-1405 */
    for (int k = 0; k < nj; ++k)
      if (e->energy[0][joiners[k]] <= 0 && !e->termd[0][joiners[k]]) handle_starvation(e, 0, joiners[k]);
  { /* prey termination (STAG:1419-1440); its observation is overwritten at STAG:551 */
    const int j = e->list_index[1][prey];
    if (!e->has_rew[j]) { e->rew[j] = 0.0; e->has_rew[j] = 1; }
    const double pen = death_penalty(c, 1, prey);
    if (pen != 0.0) e->rew[j] += pen;
    e->term[j] = 1; e->trunc[j] = 0; e->termd[1][prey] = 1;
    e->active[1] -= 1;
    *GF(e, channel_of(c, 1, prey), px, py) = 0;
    e->present[1][prey] = 0;
    remove_from_agents(e, KEY(1, prey));
    e->stats[PPG_STAT_EATEN_PREY]++;
  }
  free(snap); free(joiners); free(riders);
  return 1;
}

/* _handle_prey_engagement (STAG:1444-1500) */
static void handle_prey_engagement(env_t* e, int id) {
  const ppg_config* c = e->c;
  if (e->termd[1][id]) return;
  if (handle_team_capture(e, id)) return;
  const int px = e->x[1][id], py = e->y[1][id];
  int grass = -1;
  for (int g = 0; g < c->n_grass; ++g)
    if (e->gx[g] == px && e->gy[g] == py) { grass = g; break; }
  if (grass < 0) return;
  const int i = e->list_index[1][id];
  e->ate[i] = 1;
  const double ge = e->ge[grass];
  const int t = type_of(c, 1, id);
  double bite_size = c->bite_size_prey_t[t];
  if (!(bite_size > 0.0)) bite_size = 0.0; /* max(0.0, b) */
  double bite;
  if (t == 0) { /* mammoth leaves a rabbit bite behind (STAG:1460-1463) */
    double rb = c->bite_size_prey_t[1];
    if (!(rb > 0.0)) rb = 0.0;
    double allowed = ge - rb;
    if (!(allowed > 0.0)) allowed = 0.0;
    bite = bite_size;                       /* min(bite_size, allowed, ge): first minimal argument */
    if (allowed < bite) bite = allowed;
    if (ge < bite) bite = ge;
  } else {
    bite = bite_size < ge ? bite_size : ge; /* min(ge, bite_size) */
  }
  e->energy[1][id] += bite;
  *GF(e, channel_of(c, 1, id), px, py) = (float)e->energy[1][id];
  const double rem = ge - bite;
  e->ge[grass] = rem > 0.0 ? rem : 0.0; /* STAG:1472-1478 */
  *GF(e, CH_GRASS, px, py) = (float)e->ge[grass];
  e->stats[PPG_STAT_GRASS_EATEN]++;
}

static int occupied_by_agent(env_t* e, int x, int y) { /* `pos in set(self.agent_positions.values())` (STAG:1554,1647) */
  for (int i = 0; i < e->n_agents; ++i) {
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (e->present[s][id] && e->x[s][id] == x && e->y[s][id] == y) return 1;
  }
  return 0;
}

/* _find_available_spawn_position (STAG:1010-1044); 0 = None */
static int stag_find_spawn(env_t* e, int px, int py, int* ox, int* oy) {
  static const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
  const int G = e->G;
  for (int k = 0; k < 4; ++k) {
    const int x = px + dx[k], y = py + dy[k];
    if (x < 0 || x >= G || y < 0 || y >= G) continue;
    if (!occupied_by_agent(e, x, y) && !is_wall(e, x, y)) { *ox = x; *oy = y; return 1; } /* STAG:1026-1027 */
  }
  int n_free = 0;
  for (int cell = 0; cell < G * G; ++cell) n_free += !occupied_by_agent(e, cell / G, cell % G) && !e->wallmap[cell];
  if (n_free == 0) return 0; /* no draw (STAG:1041-1044) */
  e->stats[PPG_STAT_SPAWN_FALLBACK]++;
  int cell;
  if (take_int(e, &cell)) { *ox = cell / G; *oy = cell % G; return 1; }
  uint32_t k = ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_SPAWN, e->spawn_draws++), (uint32_t)n_free);
  for (cell = 0; cell < G * G; ++cell) /* sorted(all_positions - occupied)[k] (STAG:1039-1042) */
    if (!occupied_by_agent(e, cell / G, cell % G) && !e->wallmap[cell]) { /* STAG:1033-1039 */
      if (k == 0) { *ox = cell / G; *oy = cell % G; return 1; }
      --k;
    }
  return 0;
}

/* _handle_predator_reproduction / _handle_prey_reproduction (STAG:1502-1684) */
static void handle_reproduction(env_t* e, int s, int id) {
  const ppg_config* c = e->c;
  const int i = e->list_index[s][id];
  const int t = type_of(c, s, id);
  const double thr = s == 0 ? c->creation_threshold[0] : c->creation_threshold_prey_t[t];
  if (!(e->energy[s][id] >= thr)) return;
  if (e->next_idx_t[s][t] >= c->n_possible_t[s][t]) { e->status |= PPG_STATUS_ID_POOL_EMPTY; return; } /* STAG:1510-1523: blocked, no draw */
  if (c->cap_live[s] > 0) { /* device slot capacity (not in the reference) */
    int cnt = 0;
    for (int k = 0; k < e->n_rows; ++k) cnt += (KEY_S(e->row_key[k]) == s);
    if (cnt >= c->cap_live[s]) { e->status |= PPG_STATUS_SLOT_OVERFLOW; return; }
  }
  /* _inherit_predator_trait (STAG:1090-1097): the draws precede the spawn search (STAG:1535 before :1555) */
  double child_trait = 1.0;
  if (s == 0 && c->coop_trait_enabled) {
    child_trait = e->trait[id];
    if (c->coop_trait_mutation_std > 0.0) {
      double u, d;
      if (!take_real(e, &u)) u = ppg_draw_u01(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
      if (u < c->coop_trait_mutation_rate) {
        if (!take_real(e, &d)) d = c->coop_trait_mutation_std * ppg_draw_normal(e->seed_key, genv(e), e->episode, PPG_STREAM_TRAIT, &e->trait_draws);
        child_trait += d;
      }
    }
    child_trait = clip01(child_trait);
  }
  int sx, sy;
  if (!stag_find_spawn(e, e->x[s][id], e->y[s][id], &sx, &sy)) { /* reference: TypeError on `*None` (STAG:1566) */
    e->status |= PPG_STATUS_NO_SPAWN_CELL; /* the lockstep layer drops the birth but keeps the consumed draws */
    return;
  }
  const int child = (t ? c->n_possible_t[s][0] : 0) + e->next_idx_t[s][t]++; /* _alloc_new_id (STAG:2240-2253) */
  const int ci = e->n_rows;
  eco_ensure_rows(e, ci + 2);
  e->agents[e->n_agents++] = KEY(s, child); /* STAG:1525,1619 */
  e->row_key[e->n_rows++] = KEY(s, child);
  clear_row(e, ci);
  e->list_index[s][child] = ci;
  e->newborn[ci] = 1;
  e->age[s][child] = 0;
  e->termd[s][child] = 0;
  e->present[s][child] = 1; e->x[s][child] = (int16_t)sx; e->y[s][child] = (int16_t)sy;
  const double child_e = s == 0 ? c->initial_energy[0] : c->initial_energy_prey_t[t];
  if (s == 0) {
    e->trait[child] = child_trait;
    int f;
    if (!take_int(e, &f)) f = (int)ppg_bounded(ppg_draw_u32(e->seed_key, genv(e), e->episode, PPG_STREAM_FACING, e->facing_draws++), 8u); /* STAG:1559 */
    e->facing[child] = (int8_t)f;
    e->join[child] = 2;
    e->capture[10] += 1;
  } else {
    e->capture[11] += 1;
  }
  e->energy[s][child] = child_e;
  e->energy[s][id] -= child_e;
  *GF(e, channel_of(c, s, child), sx, sy) = (float)child_e;                       /* STAG:1566,1660 */
  *GF(e, channel_of(c, s, id), e->x[s][id], e->y[s][id]) = (float)e->energy[s][id]; /* STAG:1567,1661 */
  e->active[s] += 1;
  e->rew[ci] = 0.0; e->has_rew[ci] = 1;
  e->rew[i] = c->reproduction_reward_t[s][t]; e->has_rew[i] = 1; /* STAG:1573,1667: overwrites */
  e->repro[i] = 1;
  e->term[ci] = 0; e->trunc[ci] = 0;
  e->stats[s == 0 ? PPG_STAT_BIRTHS_PRED : PPG_STAT_BIRTHS_PREY]++;
}

/*
 * step(action_dict) (STAG:432-718).  The action dict is given in the caller's iteration order; only the movement loop
 * iterates it (STAG:805).  a_val = move | join_hunt << 8 (include/ppg.h).  Returns -1 if a key is not a possible agent.
 */
int stag_env_step(env_t* e, int n_act, const int32_t* a_s, const int32_t* a_id, const int32_t* a_val) {
  const ppg_config* c = e->c;
  e->env_flags = 0;
  const int n0 = e->n_agents; /* the ended agents strict_rllib_output leaves in self.agents are purged first (STAG:724-734) */
  eco_ensure_rows(e, n0 + 1);
  e->n_rows = n0;
  for (int i = 0; i < n0; ++i) {
    clear_row(e, i);
    e->row_key[i] = e->agents[i];
    const int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    e->list_index[s][id] = i;
    e->termd[s][id] = 0;
    if (s == 0) e->join[id] = 2; /* predator_join_intent = {} (STAG:436) */
  }
  for (int k = 0; k < n_act; ++k) /* keys of agents that are not alive are skipped, not an error (STAG:806-807) */
    if (a_s[k] < 0 || a_s[k] > 1 || a_id[k] < 0 || a_id[k] >= c->n_possible[a_s[k]]) return -1;
  e->stats[PPG_STAT_ENV_STEPS]++;
  e->stats[PPG_STAT_AGENT_STEPS] += n0;

  /* Step 1: _apply_time_step_update (STAG:720-759) over list(self.agents) */
  for (int i = 0; i < n0; ++i) {
    const int s = KEY_S(e->row_key[i]), id = KEY_ID(e->row_key[i]);
    const double loss = s == 0 ? c->energy_loss[0] : c->energy_loss_prey_t[type_of(c, 1, id)];
    e->energy[s][id] -= loss;
    *GF(e, channel_of(c, s, id), e->x[s][id], e->y[s][id]) = (float)e->energy[s][id];
    e->age[s][id] += 1;
  }
  /* Step 2: _regenerate_grass_energy (STAG:761-769) */
  for (int g = 0; g < c->n_grass; ++g) {
    const double v = e->ge[g] + c->energy_gain_grass;
    e->ge[g] = v < c->max_energy_grass ? v : c->max_energy_grass;
    *GF(e, CH_GRASS, e->gx[g], e->gy[g]) = (float)e->ge[g];
  }
  /* Step 3: _process_agent_movements (STAG:801-890), action-dict order */
  for (int k = 0; k < n_act; ++k) {
    const int s = a_s[k], id = a_id[k];
    if (!e->present[s][id] || e->termd[s][id]) continue; /* STAG:806-807 */
    int move = a_val[k] & 0xFF;
    if (s == 0) e->join[id] = (uint8_t)((a_val[k] >> PPG_STAG_JOIN_SHIFT) & 1); /* STAG:810-812 */
    const int R = c->type_action_range[type_of(c, s, id)];
    int dx = 0, dy = 0;
    if (R > 0) { /* _generate_action_map (STAG:181-190) */
      if (a_val[k] < 0 || move >= R * R) { e->status |= PPG_STATUS_BAD_ACTION; move = (R * R) / 2; } /* reference: KeyError */
      const int d = (R - 1) / 2;
      dx = move / R - d; dy = move % R - d;
    } else if (move != 0) {
      e->status |= PPG_STATUS_BAD_ACTION;
    }
    if (s == 0 && (dx != 0 || dy != 0)) { /* _update_predator_facing: from the intended move, even if blocked (STAG:933-937) */
      const int kx = (dx > 0) - (dx < 0), ky = (dy > 0) - (dy < 0);
      const int q = (kx + 1) * 3 + (ky + 1);
      e->facing[id] = (int8_t)(q < 4 ? q : q - 1);
    }
    const int ox = e->x[s][id], oy = e->y[s][id];
    int nx = clipi(ox + dx, 0, e->G - 1), ny = clipi(oy + dy, 0, e->G - 1);
    if (is_wall(e, nx, ny)) { nx = ox; ny = oy; } /* STAG:861-864 */
    else if (s == 0) { /* STAG:865-868: the `elif "predator" in agent` arm ends the chain: predators never reach the LOS test */
      if (*GF(e, CH_PRED, nx, ny) > 0) { nx = ox; ny = oy; }
    } else if (*GF(e, CH_PREY1, nx, ny) > 0 || *GF(e, CH_PREY2, nx, ny) > 0) { /* STAG:869-874 */
      nx = ox; ny = oy;
    } else {
      if (c->respect_los_for_movement && (nx != ox || ny != oy)) { /* STAG:875-889 (prey only, see above) */
        const int mx = nx - ox, my = ny - oy;
        if (abs(mx) == 1 && abs(my) == 1) { /* no corner cutting */
          if (is_wall(e, ox + mx, oy) || is_wall(e, ox, oy + my)) { nx = ox; ny = oy; }
        } else if (!los_clear(e, ox, oy, nx, ny)) { nx = ox; ny = oy; }
      }
    }
    const int ch = channel_of(c, s, id);
    *GF(e, ch, ox, oy) = 0;                         /* STAG:819,824 */
    *GF(e, ch, nx, ny) = (float)e->energy[s][id];   /* STAG:820,825 */
    e->x[s][id] = (int16_t)nx; e->y[s][id] = (int16_t)ny;
  }
  /* Step 4a: starvation over tuple(agent_energies.items()) = self.agents order (STAG:449-453) */
  for (int i = 0; i < n0; ++i) {
    const int s = KEY_S(e->row_key[i]), id = KEY_ID(e->row_key[i]);
    if (e->present[s][id] && e->energy[s][id] <= 0) handle_starvation(e, s, id);
  }
  /* Step 4b: prey engagements over a snapshot of prey_positions (STAG:456-460) */
  {
    int32_t* snap = (int32_t*)malloc(sizeof(int32_t) * (size_t)(e->n_agents + 1));
    int ns = 0;
    for (int i = 0; i < e->n_agents; ++i)
      if (KEY_S(e->agents[i]) == 1) snap[ns++] = KEY_ID(e->agents[i]);
    for (int k = 0; k < ns; ++k)
      if (!e->termd[1][snap[k]]) handle_prey_engagement(e, snap[k]);
    free(snap);
  }
  /* Step 5 (STAG:462-481) finds nothing: every ended agent has already been removed by _remove_agent_from_state */
  /* Step 7: reproduction over snapshots, predators then prey (STAG:483-496) */
  {
    const int n_live = e->n_agents;
    int32_t* snap = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_live + 1));
    memcpy(snap, e->agents, sizeof(int32_t) * (size_t)n_live);
    for (int sp = 0; sp < 2; ++sp)
      for (int i = 0; i < n_live; ++i)
        if (KEY_S(snap[i]) == sp) handle_reproduction(e, sp, KEY_ID(snap[i]));
    free(snap);
  }
  /* Step 8: outputs (STAG:548-628) */
  int live[2] = {0, 0};
  for (int i = 0; i < e->n_agents; ++i) live[KEY_S(e->agents[i])]++;
  const int episode_done = live[0] <= 0 || live[1] <= 0; /* len(predator_positions) / len(prey_positions) (STAG:584-586) */
  for (int i = 0; i < e->n_rows; ++i) {
    const int s = KEY_S(e->row_key[i]), id = KEY_ID(e->row_key[i]);
    double* o = e->obs + (size_t)i * e->max_row_elems;
    e->has_obs[i] = 1;
    if (e->term[i] == 1) { /* ended: all-zero observation (STAG:596-612) */
      memset(o, 0, sizeof(double) * (size_t)e->row_elems[s]);
      /* without strict_rllib_output the rewards dict is rebuilt from the live ids only (STAG:577) and the ended
         agents re-enter it with 0.0 (STAG:609): their death penalty is lost */
      if (!c->strict_rllib_output) e->rew[i] = 0.0;
      continue;
    }
    stag_get_observation(e, s, id, o); /* STAG:551 */
    if (!e->has_rew[i]) { e->rew[i] = 0.0; e->has_rew[i] = 1; }
    e->term[i] = 0;
    e->trunc[i] = episode_done ? 1 : 0; /* STAG:587-591: survivors of the last step are truncated, not terminated */
  }
  e->all_term = (uint8_t)episode_done; /* STAG:593-594 */
  e->all_trunc = 0;
  e->current_step += 1; /* STAG:657 */
  if (e->current_step >= c->max_steps) { /* STAG:659-716 */
    for (int i = 0; i < e->n_rows; ++i)
      if (e->term[i] != 1) { e->trunc[i] = 1; e->term[i] = 0; }
    e->all_trunc = 1; e->all_term = 0;
  }
  if (e->all_term) e->env_flags |= PPG_ENV_TERMINATED;
  if (e->all_trunc) e->env_flags |= PPG_ENV_TRUNCATED;
  e->cur_num[0] = live[0]; e->cur_num[1] = live[1];
  return 0;
}
