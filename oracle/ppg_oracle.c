/*
 * ppg_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see ppg_oracle.h).
 *
 * Sequential restatement of the reference Python env for the BASE family:
 *   BASE = predpreygrass/non_evolutionary/base_environment/predpreygrass_rllib_env.py
 *   DENSE = .../project_reward_shaping/base_environment_dense_rewards/predpreygrass_rllib_env.py
 *   ADD  = .../project_reward_shaping/base_environment_dense_rewards_additive/predpreygrass_rllib_env.py
 *   KICK = .../project_reward_shaping/base_environment_sparse_rewards_plus_kickback/predpreygrass_rllib_env.py
 * Every function cites the lines it follows.  The grid is kept persistently in float64 exactly
 * like `self.grid_world_state` (BASE:124) — the CUDA path rebuilds it per step instead, and the
 * parity tests are what shows that the two agree.
 *
 * Pinned against golden trajectories recorded from the unmodified reference
 * (tests/golden/make_golden.py -> tests/golden/base_*.npz, tests/test_oracle_golden.py).
 */
#include "ppg_oracle_int.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ppg_philox.h"

static int cmp_lex(const void* a, const void* b) {
  char sa[16], sb[16];
  snprintf(sa, sizeof sa, "%d", *(const int32_t*)a);
  snprintf(sb, sizeof sb, "%d", *(const int32_t*)b);
  return strcmp(sa, sb);
}

/* rank of str(i) among str(0..n-1): the order list.sort() gives f"{species}_{i}" (BASE:468) */
static int32_t* build_lexrank(int n) {
  int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t* rank = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) ids[i] = i;
  qsort(ids, (size_t)n, sizeof(int32_t), cmp_lex);
  for (int i = 0; i < n; ++i) rank[ids[i]] = i;
  free(ids);
  return rank;
}

static inline double* G_AT(env_t* e, int ch, int x, int y) {
  return &e->grid[((size_t)ch * e->G + x) * e->G + y];
}

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* _obs_clip + _get_observation (BASE:511-539) */
static void get_observation(env_t* e, int s, int id, double* out) {
  const int R = e->c->obs_range[s], G = e->G, C = e->C;
  const int off = (R - 1) / 2;
  const int xp = e->x[s][id], yp = e->y[s][id];
  const int xld = xp - off, xhd = xp + off, yld = yp - off, yhd = yp + off;
  const int xlo = clipi(xld, 0, G - 1), xhi = clipi(xhd, 0, G - 1);
  const int ylo = clipi(yld, 0, G - 1), yhi = clipi(yhd, 0, G - 1);
  const int xolo = abs(clipi(xld, -off, 0)), yolo = abs(clipi(yld, -off, 0));
  const int xohi = xolo + (xhi - xlo), yohi = yolo + (yhi - ylo);
  memset(out, 0, sizeof(double) * (size_t)(C * R * R));
  for (int i = 0; i < R * R; ++i) out[i] = 1.0; /* observation[0].fill(1) BASE:522 */
  for (int i = xolo; i <= xohi; ++i)
    for (int j = yolo; j <= yohi; ++j) {
      out[i * R + j] = 0.0; /* BASE:523 */
      for (int ch = 1; ch < C; ++ch)
        out[(ch * R + i) * R + j] = *G_AT(e, ch, xlo + (i - xolo), ylo + (j - yolo)); /* BASE:524 */
    }
}

static void env_alloc(env_t* e, const ppg_config* c, int env_index, int32_t* const lexrank[2]) {
  memset(e, 0, sizeof *e);
  e->c = c;
  e->G = c->grid_size;
  e->C = c->num_obs_channels;
  e->env_index = env_index;
  for (int s = 0; s < 2; ++s) {
    size_t n = (size_t)c->n_possible[s];
    e->present[s] = (uint8_t*)calloc(n, 1);
    e->x[s] = (int16_t*)calloc(n, sizeof(int16_t));
    e->y[s] = (int16_t*)calloc(n, sizeof(int16_t));
    e->energy[s] = (double*)calloc(n, sizeof(double));
    e->parent[s] = (int32_t*)malloc(n * sizeof(int32_t));
    e->list_index[s] = (int32_t*)malloc(n * sizeof(int32_t));
    e->row_elems[s] = c->num_obs_channels * c->obs_range[s] * c->obs_range[s];
    e->lexrank[s] = lexrank[s];
  }
  e->max_row_elems = e->row_elems[0] > e->row_elems[1] ? e->row_elems[0] : e->row_elems[1];
  int tot = c->n_possible[0] + c->n_possible[1];
  e->agents = (int32_t*)malloc(sizeof(int32_t) * (size_t)tot);
  e->pending = (int32_t*)malloc(sizeof(int32_t) * (size_t)tot);
  e->grid = (double*)calloc((size_t)e->C * e->G * e->G, sizeof(double));
  e->gx = (int16_t*)calloc((size_t)c->n_grass, sizeof(int16_t));
  e->gy = (int16_t*)calloc((size_t)c->n_grass, sizeof(int16_t));
  e->ge = (double*)calloc((size_t)c->n_grass, sizeof(double));
  e->cap_rows = 64;
  e->obs = NULL;
  e->seed_key = c->seed;
  e->idle = 1; /* not reset yet */
  if (c->variant == PPG_VARIANT_ECO) eco_env_alloc(e);
  if (c->variant == PPG_VARIANT_STAG) stag_env_alloc(e);
}

static void env_free(env_t* e) {
  if (e->c->variant == PPG_VARIANT_ECO) eco_env_free(e);
  if (e->c->variant == PPG_VARIANT_STAG) stag_env_free(e);
  free(e->carcass); free(e->born_obs); free(e->frozen);
  for (int s = 0; s < 2; ++s) {
    free(e->present[s]); free(e->x[s]); free(e->y[s]); free(e->energy[s]); free(e->parent[s]);
    free(e->list_index[s]);
  }
  free(e->agents); free(e->pending); free(e->grid); free(e->gx); free(e->gy); free(e->ge);
  free(e->obs); free(e->rew); free(e->has_rew); free(e->term); free(e->trunc); free(e->has_obs);
  free(e->ate); free(e->repro); free(e->newborn); free(e->e_before); free(e->bonus);
}

static void ensure_rows(env_t* e, int need) {
  if (e->obs && need <= e->cap_rows) return;
  while (e->cap_rows < need) e->cap_rows *= 2;
  size_t n = (size_t)e->cap_rows;
  e->obs = (double*)realloc(e->obs, n * (size_t)e->max_row_elems * sizeof(double));
  e->rew = (double*)realloc(e->rew, n * sizeof(double));
  e->has_rew = (int8_t*)realloc(e->has_rew, n);
  e->term = (int8_t*)realloc(e->term, n);
  e->trunc = (int8_t*)realloc(e->trunc, n);
  e->has_obs = (int8_t*)realloc(e->has_obs, n);
  e->ate = (uint8_t*)realloc(e->ate, n);
  e->repro = (uint8_t*)realloc(e->repro, n);
  e->newborn = (uint8_t*)realloc(e->newborn, n);
  e->e_before = (double*)realloc(e->e_before, n * sizeof(double));
  e->bonus = (double*)realloc(e->bonus, n * sizeof(double));
  e->carcass = (uint8_t*)realloc(e->carcass, n);
  e->born_obs = (uint8_t*)realloc(e->born_obs, n);
  e->frozen = (uint8_t*)realloc(e->frozen, n);
}

void eco_ensure_rows(env_t* e, int need) { ensure_rows(e, need); }

static void clear_row(env_t* e, int i) {
  e->rew[i] = 0.0; e->has_rew[i] = 0; e->term[i] = -1; e->trunc[i] = -1; e->has_obs[i] = 0;
  e->ate[i] = 0; e->repro[i] = 0; e->newborn[i] = 0; e->e_before[i] = 0.0; e->bonus[i] = 0.0;
}

/* reset() from explicit unique cells in the order predators, prey, grass (BASE:179-217) */
static void env_reset_cells(env_t* e, const int32_t* cells) {
  const ppg_config* c = e->c;
  const int G = e->G;
  e->current_step = 0;                                             /* BASE:134 */
  memset(e->grid, 0, sizeof(double) * (size_t)e->C * G * G);       /* BASE:138 */
  for (int s = 0; s < 2; ++s) {
    memset(e->present[s], 0, (size_t)c->n_possible[s]);            /* BASE:146-147 */
    for (int i = 0; i < c->n_possible[s]; ++i) e->parent[s][i] = -1; /* KICK:178 */
  }
  e->n_agents = 0;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < c->n_initial[s]; ++i) e->agents[e->n_agents++] = KEY(s, i); /* BASE:143-145 */
  e->n_pending = 0;                                                /* BASE:152 */
  e->next_idx[0] = c->n_initial[0];                                /* BASE:153-154 */
  e->next_idx[1] = c->n_initial[1];
  int k = 0;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < c->n_initial[s]; ++i, ++k) {               /* BASE:190-200 */
      int cx = cells[k] / G, cy = cells[k] % G;
      e->present[s][i] = 1; e->x[s][i] = (int16_t)cx; e->y[s][i] = (int16_t)cy;
      e->energy[s][i] = c->initial_energy[s];
      *G_AT(e, 1 + s, cx, cy) = c->initial_energy[s];
    }
  for (int g = 0; g < c->n_grass; ++g, ++k) {                      /* BASE:203-208 */
    e->gx[g] = (int16_t)(cells[k] / G); e->gy[g] = (int16_t)(cells[k] % G);
    e->ge[g] = c->initial_energy_grass;
    *G_AT(e, 3, e->gx[g], e->gy[g]) = c->initial_energy_grass;
  }
  e->cur_num[0] = c->n_initial[0];                                 /* BASE:210-211 */
  e->cur_num[1] = c->n_initial[1];
  /* observations = {agent: obs for agent in self.agents} (BASE:215) */
  ensure_rows(e, e->n_agents);
  for (int i = 0; i < e->n_agents; ++i) {
    clear_row(e, i);
    int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    get_observation(e, s, id, e->obs + (size_t)i * e->max_row_elems);
    e->has_obs[i] = 1; e->has_rew[i] = 1; e->term[i] = 0; e->trunc[i] = 0;
  }
  e->n_rows = e->n_agents;
  e->all_term = e->all_trunc = 0;
  e->env_flags = PPG_ENV_RESET;
  e->needs_reset = 0; e->idle = 0; e->status = 0;
  e->spawn_draws = 0;
}

/* normal-mode placement: Philox rejection draws until enough unique cells (same law as BASE:156-177) */
static void env_reset_auto(env_t* e) {
  const ppg_config* c = e->c;
  if (c->variant == PPG_VARIANT_ECO) { eco_env_reset_auto(e); return; }
  if (c->variant == PPG_VARIANT_STAG) { stag_env_reset_auto(e); return; }
  const int n_total = c->n_initial[0] + c->n_initial[1] + c->n_grass;
  const int ncell = e->G * e->G;
  int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_total);
  e->episode += 1;
  uint8_t sticky = 0;
  if (e->tape_cells && e->tape_pos + n_total <= e->tape_end) {
    memcpy(cells, e->tape_cells + e->tape_pos, sizeof(int32_t) * (size_t)n_total);
    e->tape_pos += n_total;
  } else {
    if (e->tape_cells) sticky = PPG_STATUS_TAPE_EXHAUSTED;
    uint8_t* taken = (uint8_t*)calloc((size_t)ncell, 1);
    int n = 0;
    for (uint32_t idx = 0; n < n_total; ++idx) {
      uint32_t cell = ppg_bounded(ppg_draw_u32(e->seed_key, (uint32_t)(e->env_index + e->c->env_index_base), e->episode,
                                               PPG_STREAM_PLACEMENT, idx), (uint32_t)ncell);
      if (!taken[cell]) { taken[cell] = 1; cells[n++] = (int32_t)cell; }
    }
    free(taken);
  }
  env_reset_cells(e, cells);
  e->status |= sticky;
  free(cells);
}

/* _get_move (BASE:495-509) */
static void get_move(env_t* e, int s, int id, int action, int* nx, int* ny) {
  const int ch = 1 + s;
  int dx = action / 3 - 1, dy = action % 3 - 1; /* action_to_move_tuple BASE:96-106 */
  int x = clipi(e->x[s][id] + dx, 0, e->G - 1), y = clipi(e->y[s][id] + dy, 0, e->G - 1);
  if (*G_AT(e, ch, x, y) > 0) { x = e->x[s][id]; y = e->y[s][id]; }
  *nx = x; *ny = y;
}

static int occupied_by_agent(env_t* e, int x, int y) {
  /* `pos in set(self.agent_positions.values())` (BASE:399,754) */
  for (int i = 0; i < e->n_agents; ++i) {
    int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (e->present[s][id] && e->x[s][id] == x && e->y[s][id] == y) return 1;
  }
  return 0;
}

/* _find_available_spawn_position (BASE:738-766); returns 0 if no cell */
static int find_spawn(env_t* e, int px, int py, int* ox, int* oy) {
  static const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1}; /* BASE:749 */
  const int G = e->G;
  for (int k = 0; k < 4; ++k) {
    int x = px + dx[k], y = py + dy[k];
    if (x < 0 || x >= G || y < 0 || y >= G) continue;
    if (!occupied_by_agent(e, x, y)) { *ox = x; *oy = y; return 1; }
  }
  /* fallback: uniformly random free cell (BASE:760-764) — from the tape, else Philox */
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
  uint32_t k = ppg_bounded(ppg_draw_u32(e->seed_key, (uint32_t)(e->env_index + e->c->env_index_base), e->episode,
                                        PPG_STREAM_SPAWN, e->spawn_draws++), (uint32_t)n_free);
  for (int cell = 0; cell < G * G; ++cell)
    if (!occupied_by_agent(e, cell / G, cell % G)) {
      if (k == 0) { *ox = cell / G; *oy = cell % G; return 1; }
      --k;
    }
  return 0;
}

static int cmp_agents_lex(const void* a, const void* b, void* ctx) {
  env_t* e = (env_t*)ctx;
  int32_t ka = *(const int32_t*)a, kb = *(const int32_t*)b;
  if (KEY_S(ka) != KEY_S(kb)) return KEY_S(ka) - KEY_S(kb); /* "predator_*" < "prey_*" */
  return e->lexrank[KEY_S(ka)][KEY_ID(ka)] - e->lexrank[KEY_S(kb)][KEY_ID(kb)];
}

/*
 * step(action_dict) (BASE:219-473), with the reward variants of DENSE/ADD/KICK.
 * The action dict is given in the caller's iteration order.  `lockstep` selects the batched
 * convention for max_steps (truncation reported on the step that reaches it, no extra call).
 */
static int env_step(env_t* e, int n_act, const int32_t* a_s, const int32_t* a_id, const int32_t* a_val,
                    int lockstep) {
  const ppg_config* c = e->c;
  if (c->variant == PPG_VARIANT_ECO) return eco_env_step(e, n_act, a_s, a_id, a_val);
  if (c->variant == PPG_VARIANT_STAG) return stag_env_step(e, n_act, a_s, a_id, a_val);
  const int mode = c->reward_mode;
  const int dense = (mode == PPG_REWARD_DENSE || mode == PPG_REWARD_DENSE_ADDITIVE);
  e->env_flags = 0;

  /* BASE:222-225 drop last step's terminated agents from self.agents */
  if (e->n_pending) {
    int w = 0;
    for (int i = 0; i < e->n_agents; ++i) {
      int drop = 0;
      for (int j = 0; j < e->n_pending; ++j) drop |= (e->pending[j] == e->agents[i]);
      if (!drop) e->agents[w++] = e->agents[i];
    }
    e->n_agents = w;
    e->n_pending = 0;
  }
  ensure_rows(e, e->n_agents + 1);
  for (int i = 0; i < e->n_agents; ++i) {
    clear_row(e, i);
    e->list_index[KEY_S(e->agents[i])][KEY_ID(e->agents[i])] = i;
  }

  /* step 0: truncation on an extra call (BASE:228-238) */
  if (!lockstep && e->current_step >= c->max_steps) {
    for (int i = 0; i < e->n_agents; ++i) {
      int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
      get_observation(e, s, id, e->obs + (size_t)i * e->max_row_elems);
      e->has_obs[i] = 1; e->rew[i] = 0.0; e->has_rew[i] = 1; e->trunc[i] = 1; e->term[i] = 0;
    }
    e->n_rows = e->n_agents;
    e->all_trunc = 1; e->all_term = 0;
    e->env_flags = PPG_ENV_TRUNCATED;
    return 0;
  }

  /* every key must be a live agent (BASE:246 KeyError otherwise) */
  for (int k = 0; k < n_act; ++k)
    if (a_id[k] < 0 || a_id[k] >= c->n_possible[a_s[k]] || !e->present[a_s[k]][a_id[k]]) return -1;

  e->stats[PPG_STAT_ENV_STEPS]++;
  e->stats[PPG_STAT_AGENT_STEPS] += e->n_agents;

  /* ADD:256 energy_before = dict(self.agent_energies) */
  for (int i = 0; i < e->n_agents; ++i)
    e->e_before[i] = e->energy[KEY_S(e->agents[i])][KEY_ID(e->agents[i])];

  /* Step 1: energy depletion, dict order (BASE:244-250) */
  for (int k = 0; k < n_act; ++k) {
    int s = a_s[k], id = a_id[k];
    e->energy[s][id] -= c->energy_loss[s];
    *G_AT(e, 1 + s, e->x[s][id], e->y[s][id]) = e->energy[s][id];
  }
  /* grass regrowth (BASE:252-256) */
  /* base_environment_seasonal: gain * _current_season_multiplier() (SEASON:224-234,268-271; SEASON =
   * predpreygrass/non_evolutionary/base_environment_seasonal/predpreygrass_rllib_env.py) */
  double gain = c->energy_gain_grass;
  if (c->season_length_steps > 0) gain = c->energy_gain_grass * c->season_multiplier[(e->current_step / c->season_length_steps) % 2];
  for (int g = 0; g < c->n_grass; ++g) {
    double v = e->ge[g] + gain;
    e->ge[g] = v < c->initial_energy_grass ? v : c->initial_energy_grass; /* min(a, b) */
    *G_AT(e, 3, e->gx[g], e->gy[g]) = e->ge[g];
  }

  /* Step 2: movements, dict order (BASE:259-273); move cost is 0 (BASE:475-493) */
  for (int k = 0; k < n_act; ++k) {
    int s = a_s[k], id = a_id[k], act = a_val[k];
    if (!e->present[s][id]) continue;
    if (act < 0 || act > 8) { e->status |= PPG_STATUS_BAD_ACTION; act = 4; }
    int ox = e->x[s][id], oy = e->y[s][id], nx, ny;
    get_move(e, s, id, act, &nx, &ny);
    e->x[s][id] = (int16_t)nx; e->y[s][id] = (int16_t)ny;
    *G_AT(e, 1 + s, ox, oy) = 0;
    *G_AT(e, 1 + s, nx, ny) = e->energy[s][id];
  }

  /* Step 3: removals and engagements over self.agents (BASE:279-380) */
  const int n_start = e->n_agents;
  for (int i = 0; i < n_start; ++i) {
    int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    double* obs_i = e->obs + (size_t)i * e->max_row_elems;
    if (!e->present[s][id]) continue; /* BASE:281 */
    if (e->energy[s][id] <= 0) {      /* BASE:284-301 */
      get_observation(e, s, id, obs_i); e->has_obs[i] = 1;
      e->rew[i] = dense ? (e->energy[s][id] - e->e_before[i]) : 0.0; /* BASE:288 / ADD:308 */
      e->has_rew[i] = 1; e->term[i] = 1; e->trunc[i] = 0;
      e->cur_num[s] -= 1;
      *G_AT(e, 1 + s, e->x[s][id], e->y[s][id]) = 0;
      e->present[s][id] = 0;
      e->parent[s][id] = -1; /* KICK:325 */
      e->stats[s == 0 ? PPG_STAT_STARVED_PRED : PPG_STAT_STARVED_PREY]++;
      continue;
    }
    if (s == PPG_PREDATOR) { /* BASE:302-346 */
      int px = e->x[0][id], py = e->y[0][id];
      /* first prey in agent_positions insertion order on the same cell (BASE:305-312):
         prey are inserted in ascending id order (founders 0..n-1, then next_idx++) */
      int caught = -1;
      for (int q = 0; q < e->next_idx[1]; ++q)
        if (e->present[1][q] && e->x[1][q] == px && e->y[1][q] == py) { caught = q; break; }
      if (caught >= 0) {
        int j = e->list_index[1][caught];
        e->ate[i] = 1;                                                   /* BASE:319 */
        if (!dense) { e->rew[i] = c->reward_predator_catch_prey; e->has_rew[i] = 1; } /* BASE:322 */
        e->energy[0][id] += e->energy[1][caught];                        /* BASE:324 */
        *G_AT(e, 1, px, py) = e->energy[0][id];                          /* BASE:325 */
        get_observation(e, 1, caught, e->obs + (size_t)j * e->max_row_elems); /* BASE:327 */
        e->has_obs[j] = 1;
        e->rew[j] = dense ? (0.0 - e->e_before[j]) : c->penalty_prey_caught; /* BASE:328 / ADD:346 */
        e->has_rew[j] = 1; e->term[j] = 1; e->trunc[j] = 0;              /* BASE:332-333 */
        e->cur_num[1] -= 1;
        *G_AT(e, 2, e->x[1][caught], e->y[1][caught]) = 0;               /* BASE:335 */
        e->present[1][caught] = 0;                                       /* BASE:336-338 */
        e->parent[1][caught] = -1;                                       /* KICK:364 */
        e->stats[PPG_STAT_EATEN_PREY]++;
      } else if (!dense) {
        e->rew[i] = c->reward_predator_step; e->has_rew[i] = 1;          /* BASE:341 */
      }
      get_observation(e, 0, id, obs_i); e->has_obs[i] = 1;               /* BASE:343 */
      e->term[i] = 0; e->trunc[i] = 0;
    } else { /* prey, BASE:347-380 */
      if (e->term[i] == 1) continue; /* BASE:348 (unreachable: caught prey left agent_positions) */
      int px = e->x[1][id], py = e->y[1][id];
      int grass = -1;
      for (int g = 0; g < c->n_grass; ++g)
        if (e->gx[g] == px && e->gy[g] == py) { grass = g; break; }      /* BASE:351-358 */
      if (grass >= 0) {
        e->ate[i] = 1;                                                   /* BASE:362 */
        if (!dense) { e->rew[i] = c->reward_prey_eat_grass; e->has_rew[i] = 1; } /* BASE:365 */
        e->energy[1][id] += e->ge[grass];                                /* BASE:367 */
        *G_AT(e, 2, px, py) = e->energy[1][id];
        *G_AT(e, 3, px, py) = 0;                                         /* BASE:371-372 */
        e->ge[grass] = 0;
        e->stats[PPG_STAT_GRASS_EATEN]++;
      } else if (!dense) {
        e->rew[i] = c->reward_prey_step; e->has_rew[i] = 1;              /* BASE:375 */
      }
      get_observation(e, 1, id, obs_i); e->has_obs[i] = 1;               /* BASE:377 */
      e->term[i] = 0; e->trunc[i] = 0;
    }
  }

  /* Step 4: schedule removals (BASE:383) */
  e->n_pending = 0;
  for (int i = 0; i < n_start; ++i)
    if (e->term[i] == 1) e->pending[e->n_pending++] = e->agents[i];

  /* Step 5: spawning over self.agents[:] (BASE:389-448) */
  for (int i = 0; i < n_start; ++i) {
    if (e->term[i] == 1) continue; /* in _pending_removal (BASE:390) */
    int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (!(e->energy[s][id] >= c->creation_threshold[s])) continue; /* BASE:393,422 */
    if (!(e->next_idx[s] < c->n_possible[s])) continue;            /* BASE:395,424 */
    /* device slot capacity (not in the reference): count this species' entries in self.agents */
    if (c->cap_live[s] > 0) {
      int cnt = 0;
      for (int k = 0; k < e->n_agents; ++k) cnt += (KEY_S(e->agents[k]) == s);
      if (cnt >= c->cap_live[s]) { e->status |= PPG_STATUS_SLOT_OVERFLOW; continue; }
    }
    int sx, sy;
    if (!find_spawn(e, e->x[s][id], e->y[s][id], &sx, &sy)) {
      e->status |= PPG_STATUS_NO_SPAWN_CELL; /* reference would raise (BASE:766 returns None) */
      continue;
    }
    int child = e->next_idx[s]++;                       /* BASE:396-397 */
    int ci = e->n_agents;
    ensure_rows(e, ci + 2);
    e->agents[e->n_agents++] = KEY(s, child);           /* BASE:398 */
    clear_row(e, ci);
    e->list_index[s][child] = ci;
    e->newborn[ci] = 1;
    e->present[s][child] = 1; e->x[s][child] = (int16_t)sx; e->y[s][child] = (int16_t)sy; /* BASE:401 */
    e->energy[s][child] = c->initial_energy[s];         /* BASE:403 */
    e->energy[s][id] -= c->initial_energy[s];           /* BASE:404 */
    e->parent[s][child] = id;                           /* KICK:434 */
    *G_AT(e, 1 + s, sx, sy) = c->initial_energy[s];     /* BASE:405 */
    *G_AT(e, 1 + s, e->x[s][id], e->y[s][id]) = e->energy[s][id]; /* BASE:406 */
    e->cur_num[s] += 1;
    e->rew[ci] = 0.0; e->has_rew[ci] = 1;               /* BASE:408 */
    e->repro[i] = 1;
    if (mode == PPG_REWARD_DENSE_ADDITIVE) e->bonus[i] = c->reproduction_reward[s]; /* ADD:419 */
    else if (mode != PPG_REWARD_DENSE) { e->rew[i] = c->reproduction_reward[s]; e->has_rew[i] = 1; } /* BASE:409 */
    if (mode == PPG_REWARD_SPARSE_KICKBACK) {           /* KICK:439-449 */
      int gp = e->parent[s][id];
      if (gp >= 0 && e->present[s][gp]) {
        int gi = e->list_index[s][gp];
        e->rew[gi] = (e->has_rew[gi] ? e->rew[gi] : 0.0) + c->kickback_reward[s];
        e->has_rew[gi] = 1;
      }
    }
    get_observation(e, s, child, e->obs + (size_t)ci * e->max_row_elems); /* BASE:412 */
    e->has_obs[ci] = 1; e->term[ci] = 0; e->trunc[ci] = 0;
    e->stats[s == 0 ? PPG_STAT_BIRTHS_PRED : PPG_STAT_BIRTHS_PREY]++;
  }

  /* Step 5b: dense reward = net energy delta (+ bonus) for survivors (ADD:468-471) */
  if (dense)
    for (int i = 0; i < n_start; ++i) {
      int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
      if (e->present[s][id]) { e->rew[i] = (e->energy[s][id] - e->e_before[i]) + e->bonus[i]; e->has_rew[i] = 1; }
    }

  /* Step 6: observations of everyone still present (BASE:451-453) */
  for (int i = 0; i < e->n_agents; ++i) {
    int s = KEY_S(e->agents[i]), id = KEY_ID(e->agents[i]);
    if (e->present[s][id]) { get_observation(e, s, id, e->obs + (size_t)i * e->max_row_elems); e->has_obs[i] = 1; }
  }

  e->all_term = (e->cur_num[1] <= 0 || e->cur_num[0] <= 0); /* BASE:466 */
  e->all_trunc = 0;                                         /* BASE:463 */
  e->n_rows = e->n_agents;                                  /* BASE:459-462 */

  /* self.agents.sort() (BASE:468) — rows were laid out in the pre-sort order, so sort a copy of
     the row order: keep `agents` sorted for the next call but remember the output order. */
  /* (rows are exported by the caller before the sort is visible: see export_rows) */
  e->current_step += 1; /* BASE:471 */

  if (e->all_term) e->env_flags |= PPG_ENV_TERMINATED;
  if (lockstep && !e->all_term && e->current_step >= c->max_steps) {
    /* batched convention: report BASE's extra-call truncation (BASE:228-238) on this step; its
       observations are those of this step (grid and positions do not change in between) */
    e->all_trunc = 1;
    e->env_flags |= PPG_ENV_TRUNCATED;
    for (int i = 0; i < e->n_agents; ++i)
      if (e->present[KEY_S(e->agents[i])][KEY_ID(e->agents[i])]) e->trunc[i] = 1;
  }
  return 0;
}

/* finish the call: self.agents.sort() (BASE:468), after the rows have been exported */
static void env_sort_agents(env_t* e) {
  if (e->c->variant != PPG_VARIANT_BASE) return; /* ECO and STAG never sort self.agents */
  /* insertion sort keeps it dependency-free (qsort_r is a GNU extension) */
  for (int i = 1; i < e->n_agents; ++i) {
    int32_t k = e->agents[i];
    int j = i - 1;
    while (j >= 0 && cmp_agents_lex(&e->agents[j], &k, e) > 0) { e->agents[j + 1] = e->agents[j]; --j; }
    e->agents[j + 1] = k;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* lockstep batch layer: row layout of include/ppg.h                                           */
/* ------------------------------------------------------------------------------------------ */
ppgo_batch* ppgo_create(const ppg_config* cfg, int32_t n_envs) {
  if (!cfg || cfg->struct_size != sizeof(ppg_config) || n_envs <= 0) return NULL;
  if (cfg->variant != PPG_VARIANT_BASE && cfg->variant != PPG_VARIANT_ECO && cfg->variant != PPG_VARIANT_STAG) return NULL;
  if (cfg->n_possible[0] > 65535 || cfg->n_possible[1] > 65535) return NULL;
  if (cfg->n_initial[0] + cfg->n_initial[1] + cfg->n_grass > cfg->grid_size * cfg->grid_size) return NULL; /* BASE:167 */
  ppgo_batch* b = (ppgo_batch*)calloc(1, sizeof *b);
  b->cfg = *cfg;
  if (cfg->n_walls > 0 && cfg->wall_cells) { /* the caller's wall list is copied (include/ppg.h) */
    int32_t* w = (int32_t*)malloc(sizeof(int32_t) * (size_t)cfg->n_walls);
    memcpy(w, cfg->wall_cells, sizeof(int32_t) * (size_t)cfg->n_walls);
    b->cfg.wall_cells = w;
  } else { b->cfg.wall_cells = NULL; b->cfg.n_walls = 0; }
  b->n_envs = n_envs;
  b->n_threads = 1;
  for (int s = 0; s < 2; ++s) b->lexrank[s] = build_lexrank(cfg->n_possible[s]);
  b->envs = (env_t*)calloc((size_t)n_envs, sizeof(env_t));
  for (int e = 0; e < n_envs; ++e) env_alloc(&b->envs[e], &b->cfg, e, b->lexrank);
  for (int s = 0; s < 2; ++s) {
    int cap = cfg->cap_live[s] > 0 ? cfg->cap_live[s] : cfg->n_possible[s];
    b->cap[s] = (int64_t)n_envs * cap;
    size_t n = (size_t)b->cap[s];
    int elems = (cfg->num_obs_channels + (cfg->variant == PPG_VARIANT_ECO && cfg->include_speed_in_obs ? 1 : 0)) * cfg->obs_range[s] * cfg->obs_range[s];
    b->out.f.obs[s] = (float*)malloc(n * (size_t)elems * sizeof(float));
    b->out.obs64[s] = (double*)malloc(n * (size_t)elems * sizeof(double));
    b->out.f.row_env[s] = (int32_t*)malloc(n * sizeof(int32_t));
    b->out.f.row_agent[s] = (int32_t*)malloc(n * sizeof(int32_t));
    b->out.f.reward[s] = (float*)malloc(n * sizeof(float));
    b->out.reward64[s] = (double*)malloc(n * sizeof(double));
    b->out.f.flags[s] = (uint8_t*)malloc(n);
    b->out.f.old_off[s] = (int32_t*)calloc((size_t)n_envs + 1, sizeof(int32_t));
    b->out.f.new_off[s] = (int32_t*)calloc((size_t)n_envs + 1, sizeof(int32_t));
    b->out.f.new_cnt[s] = (int32_t*)calloc((size_t)n_envs + 1, sizeof(int32_t));
    b->out.f.row_capacity[s] = b->cap[s];
    b->out.f.obs_row_elems[s] = elems;
    b->prev_row[s] = (int32_t*)malloc((size_t)n_envs * (size_t)cfg->n_possible[s] * sizeof(int32_t));
  }
  b->out.f.n_rows = b->n_rows;
  b->out.f.env_flags = (uint8_t*)calloc((size_t)n_envs, 1);
  b->out.f.env_status = (uint8_t*)calloc((size_t)n_envs, 1);
  b->out.f.env_step = (int32_t*)calloc((size_t)n_envs, sizeof(int32_t));
  b->out.f.env_count = (int32_t*)calloc((size_t)n_envs * 2, sizeof(int32_t));
  b->out.f.n_envs = n_envs;
  return b;
}

void ppgo_destroy(ppgo_batch* b) {
  if (!b) return;
  for (int e = 0; e < b->n_envs; ++e) env_free(&b->envs[e]);
  free(b->envs);
  for (int s = 0; s < 2; ++s) {
    free(b->lexrank[s]); free(b->out.f.obs[s]); free(b->out.obs64[s]); free(b->out.f.row_env[s]);
    free(b->out.f.row_agent[s]); free(b->out.f.reward[s]); free(b->out.reward64[s]);
    free(b->out.f.flags[s]); free(b->out.f.old_off[s]); free(b->out.f.new_off[s]); free(b->out.f.new_cnt[s]); free(b->prev_row[s]);
  }
  free(b->out.f.env_flags); free(b->out.f.env_status); free(b->out.f.env_step); free(b->out.f.env_count);
  free(b->tape_cells); free(b->tape_off); free(b->tape_reals); free(b->tape_real_off);
  free((void*)b->cfg.wall_cells);
  free(b);
}

void ppgo_set_threads(ppgo_batch* b, int32_t n) { b->n_threads = n < 1 ? 1 : n; }

int ppgo_load_tape(ppgo_batch* b, const ppg_tape* t) {
  free(b->tape_cells); free(b->tape_off); free(b->tape_reals); free(b->tape_real_off);
  b->tape_cells = NULL; b->tape_off = NULL; b->tape_reals = NULL; b->tape_real_off = NULL; b->has_tape = 0;
  for (int e = 0; e < b->n_envs; ++e) {
    b->envs[e].tape_cells = NULL; b->envs[e].tape_pos = b->envs[e].tape_end = 0;
    b->envs[e].tape_reals = NULL; b->envs[e].real_pos = b->envs[e].real_end = 0;
  }
  if (t && t->reals && t->real_off) {
    int64_t total = t->real_off[b->n_envs];
    b->tape_reals = (double*)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
    b->tape_real_off = (int64_t*)malloc(sizeof(int64_t) * ((size_t)b->n_envs + 1));
    memcpy(b->tape_reals, t->reals, sizeof(double) * (size_t)total);
    memcpy(b->tape_real_off, t->real_off, sizeof(int64_t) * ((size_t)b->n_envs + 1));
    for (int e = 0; e < b->n_envs; ++e) {
      b->envs[e].tape_reals = b->tape_reals;
      b->envs[e].real_pos = b->tape_real_off[e];
      b->envs[e].real_end = b->tape_real_off[e + 1];
    }
  }
  if (!t || !t->cells || !t->cell_off) return PPG_OK;
  int64_t total = t->cell_off[b->n_envs];
  b->tape_cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
  b->tape_off = (int64_t*)malloc(sizeof(int64_t) * ((size_t)b->n_envs + 1));
  memcpy(b->tape_cells, t->cells, sizeof(int32_t) * (size_t)total);
  memcpy(b->tape_off, t->cell_off, sizeof(int64_t) * ((size_t)b->n_envs + 1));
  for (int e = 0; e < b->n_envs; ++e) {
    b->envs[e].tape_cells = b->tape_cells;
    b->envs[e].tape_pos = b->tape_off[e];
    b->envs[e].tape_end = b->tape_off[e + 1];
  }
  b->has_tape = 1;
  return PPG_OK;
}

/* lay the per-env rows out as [old rows by env][new rows by env] per species */
static void export_rows(ppgo_batch* b) {
  const ppg_config* c = &b->cfg;
  int32_t n_old[2] = {0, 0}, n_new[2] = {0, 0};
  for (int e = 0; e < b->n_envs; ++e) {
    env_t* v = &b->envs[e];
    for (int s = 0; s < 2; ++s) { b->out.f.old_off[s][e] = n_old[s]; }
    const int32_t* keys = v->row_key ? v->row_key : v->agents;
    for (int i = 0; i < v->n_rows; ++i)
      if (v->has_obs[i] && !v->newborn[i]) n_old[KEY_S(keys[i])]++;
  }
  for (int s = 0; s < 2; ++s) b->out.f.old_off[s][b->n_envs] = n_old[s];
  for (int e = 0; e < b->n_envs; ++e) {
    env_t* v = &b->envs[e];
    for (int s = 0; s < 2; ++s) b->out.f.new_off[s][e] = n_old[s] + n_new[s];
    const int32_t* keys = v->row_key ? v->row_key : v->agents;
    for (int i = 0; i < v->n_rows; ++i)
      if (v->has_obs[i] && v->newborn[i]) n_new[KEY_S(keys[i])]++;
  }
  for (int s = 0; s < 2; ++s) b->out.f.new_off[s][b->n_envs] = n_old[s] + n_new[s];
  for (int s = 0; s < 2; ++s) /* (start, count) form of include/ppg.h: start is 0 where the env has no newborn rows */
    for (int e = 0; e < b->n_envs; ++e) b->out.f.new_cnt[s][e] = b->out.f.new_off[s][e + 1] - b->out.f.new_off[s][e];
  b->n_rows[0] = n_old[0]; b->n_rows[1] = n_old[1]; b->n_rows[2] = n_new[0]; b->n_rows[3] = n_new[1];

  for (int e = 0; e < b->n_envs; ++e) {
    env_t* v = &b->envs[e];
    int32_t po[2] = {b->out.f.old_off[0][e], b->out.f.old_off[1][e]};
    int32_t pn[2] = {b->out.f.new_off[0][e], b->out.f.new_off[1][e]};
    const int32_t* keys = v->row_key ? v->row_key : v->agents;
    for (int i = 0; i < v->n_rows; ++i) {
      if (!v->has_obs[i]) continue;
      int s = KEY_S(keys[i]), id = KEY_ID(keys[i]);
      int32_t row = v->newborn[i] ? pn[s]++ : po[s]++;
      int elems = v->row_elems[s];
      const double* src = v->obs + (size_t)i * v->max_row_elems;
      double* d64 = b->out.obs64[s] + (size_t)row * elems;
      float* d32 = b->out.f.obs[s] + (size_t)row * elems;
      for (int k = 0; k < elems; ++k) { d64[k] = src[k]; d32[k] = (float)src[k]; }
      b->out.f.row_env[s][row] = e;
      b->out.f.row_agent[s][row] = id;
      b->out.reward64[s][row] = v->rew[i];
      b->out.f.reward[s][row] = (float)v->rew[i];
      uint8_t fl = 0;
      if (v->term[i] == 1) fl |= PPG_ROW_TERMINATED;
      if (v->trunc[i] == 1) fl |= PPG_ROW_TRUNCATED;
      if (v->newborn[i]) fl |= PPG_ROW_NEWBORN;
      if (v->env_flags & PPG_ENV_RESET) fl |= PPG_ROW_FOUNDER;
      if (v->ate[i]) fl |= PPG_ROW_ATE;
      if (v->repro[i]) fl |= PPG_ROW_REPRODUCED;
      if (v->row_key && v->carcass[i]) fl |= PPG_ROW_CARCASS;
      if (v->row_key && c->trait_mode == PPG_TRAIT_CADENCE && v->frozen[i]) fl |= PPG_ROW_FROZEN;
      b->out.f.flags[s][row] = fl;
      b->prev_row[s][(size_t)e * c->n_possible[s] + id] = row;
    }
    b->out.f.env_flags[e] = v->env_flags;
    b->out.f.env_status[e] = v->status;
    b->out.f.env_step[e] = v->current_step;
    b->out.f.env_count[2 * e] = v->cur_num[0];
    b->out.f.env_count[2 * e + 1] = v->cur_num[1];
  }
  for (int s = 0; s < 2; ++s)
    for (int e = 0; e < b->n_envs; ++e)
      if (b->out.f.new_cnt[s][e] == 0) b->out.f.new_off[s][e] = 0;
}

int ppgo_reset(ppgo_batch* b, const uint64_t* seeds, const uint8_t* mask) {
  if (mask) {
    /* partial reset: scheduled, performed by the next ppgo_step (the other envs keep their rows) */
    for (int e = 0; e < b->n_envs; ++e)
      if (mask[e]) {
        if (seeds) b->envs[e].seed_key = seeds[e];
        b->envs[e].needs_reset = 1;
        b->envs[e].idle = 0;
      }
    return PPG_OK;
  }
  for (int e = 0; e < b->n_envs; ++e) {
    if (seeds) b->envs[e].seed_key = seeds[e];
    env_reset_auto(&b->envs[e]);
  }
  export_rows(b);
  b->calls++;
  return PPG_OK;
}

typedef struct job { ppgo_batch* b; const int32_t* act[2]; int e0, e1; int rc; } job;

static void step_range(job* j) {
  ppgo_batch* b = j->b;
  const ppg_config* c = &b->cfg;
  int cap = 0;
  int32_t *as = NULL, *aid = NULL, *av = NULL;
  for (int e = j->e0; e < j->e1; ++e) {
    env_t* v = &b->envs[e];
    if (v->idle) { v->n_rows = 0; v->env_flags = PPG_ENV_IDLE; continue; }
    if (v->needs_reset) { env_reset_auto(v); continue; }
    /* the action dict in the iteration order of the previous observation dict:
       rows of the previous call, terminated ones left out */
    int n_prev = v->n_rows;
    if (n_prev > cap) {
      cap = n_prev * 2;
      as = (int32_t*)realloc(as, sizeof(int32_t) * (size_t)cap);
      aid = (int32_t*)realloc(aid, sizeof(int32_t) * (size_t)cap);
      av = (int32_t*)realloc(av, sizeof(int32_t) * (size_t)cap);
    }
    /* v->agents has been sorted after the export; the previous dict order is recovered from the
       row numbers: old rows ascending, then newborn rows ascending, per species */
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)     /* pass 0: old rows, pass 1: newborn rows */
      for (int s = 0; s < 2; ++s) {
        const int32_t r0 = pass == 0 ? b->out.f.old_off[s][e] : b->out.f.new_off[s][e];
        const int32_t r1 = pass == 0 ? b->out.f.old_off[s][e + 1] : r0 + b->out.f.new_cnt[s][e];
        for (int32_t row = r0; row < r1; ++row) {
          if (b->out.f.flags[s][row] & PPG_ROW_TERMINATED) continue;
          as[n] = s; aid[n] = b->out.f.row_agent[s][row]; av[n] = j->act[s][row]; ++n;
        }
      }
    if (env_step(v, n, as, aid, av, 1) != 0) j->rc = PPG_ERR_STATE;
  }
  free(as); free(aid); free(av);
}

static void* step_thread(void* p) { step_range((job*)p); return NULL; }

int ppgo_step(ppgo_batch* b, const int32_t* actions_pred, const int32_t* actions_prey) {
  int nt = b->n_threads;
  if (nt > b->n_envs) nt = b->n_envs;
  int rc = PPG_OK;
  if (nt <= 1) {
    job j = {b, {actions_pred, actions_prey}, 0, b->n_envs, PPG_OK};
    step_range(&j);
    rc = j.rc;
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    job* js = (job*)malloc(sizeof(job) * (size_t)nt);
    for (int t = 0; t < nt; ++t) {
      js[t].b = b; js[t].act[0] = actions_pred; js[t].act[1] = actions_prey;
      js[t].e0 = (int)((int64_t)b->n_envs * t / nt);
      js[t].e1 = (int)((int64_t)b->n_envs * (t + 1) / nt);
      js[t].rc = PPG_OK;
      pthread_create(&th[t], NULL, step_thread, &js[t]);
    }
    for (int t = 0; t < nt; ++t) { pthread_join(th[t], NULL); if (js[t].rc) rc = js[t].rc; }
    free(th); free(js);
  }
  export_rows(b);
  for (int e = 0; e < b->n_envs; ++e) {
    env_t* v = &b->envs[e];
    if (v->env_flags & (PPG_ENV_RESET | PPG_ENV_IDLE)) continue;
    env_sort_agents(v); /* BASE:468 */
    if (v->env_flags & (PPG_ENV_TERMINATED | PPG_ENV_TRUNCATED)) {
      v->stats[PPG_STAT_EPISODES]++;
      v->stats[PPG_STAT_EPISODE_STEPS] += v->current_step;
      if (v->env_flags & PPG_ENV_TRUNCATED) v->stats[PPG_STAT_TRUNCATED]++;
      if (b->cfg.autoreset) v->needs_reset = 1; else v->idle = 1;
    }
  }
  b->calls++;
  return rc;
}

int ppgo_random_actions(ppgo_batch* b, uint64_t seed, int32_t* actions_pred, int32_t* actions_prey) {
  int32_t* act[2] = {actions_pred, actions_prey};
  for (int s = 0; s < 2; ++s) {
    int32_t n = b->n_rows[s] + b->n_rows[2 + s];
    for (int32_t row = 0; row < n; ++row) {
      uint32_t env = (uint32_t)(b->out.f.row_env[s][row] + b->cfg.env_index_base), id = (uint32_t)b->out.f.row_agent[s][row];
      uint32_t r = ppg_draw_u32(seed, env, (uint32_t)b->calls, PPG_STREAM_ACTION + 8u * (uint32_t)s, id);
      if (b->cfg.variant == PPG_VARIANT_STAG) { /* MultiDiscrete([n_moves, 2]) for predators, Discrete(n_moves) for prey (STAG:1802-1816) */
        const int R = b->cfg.type_action_range[(int)id >= b->cfg.n_possible_t[s][0]];
        const uint32_t mv = ppg_bounded(r, (uint32_t)(R > 0 ? R * R : 1));
        const uint32_t jn = s == 0 ? (ppg_draw_u32(seed, env, (uint32_t)b->calls, PPG_STREAM_ACTION + 16u, id) & 1u) : 0u;
        act[s][row] = (int32_t)(mv | (jn << PPG_STAG_JOIN_SHIFT));
        continue;
      }
      act[s][row] = (int32_t)ppg_bounded(r, (uint32_t)(b->cfg.variant == PPG_VARIANT_ECO ? b->cfg.action_range * b->cfg.action_range : 9));
    }
  }
  return PPG_OK;
}

int ppgo_get_buffers(ppgo_batch* b, ppgo_buffers* out) { *out = b->out; return PPG_OK; }

int ppgo_stats(ppgo_batch* b, int64_t* out) {
  memset(out, 0, sizeof(int64_t) * PPG_N_STATS);
  for (int e = 0; e < b->n_envs; ++e) {
    for (int k = 0; k < PPG_N_STATS; ++k) out[k] += b->envs[e].stats[k];
    out[PPG_STAT_STATUS_ENVS] += b->envs[e].status != 0;
  }
  return PPG_OK;
}

int ppgo_read_env(ppgo_batch* b, int32_t env, int32_t* n_live, int32_t* ids_pred, int32_t* xy_pred,
                  double* energy_pred, int32_t* ids_prey, int32_t* xy_prey, double* energy_prey,
                  int32_t* xy_grass, double* energy_grass) {
  if (env < 0 || env >= b->n_envs) return PPG_ERR_INVALID;
  env_t* v = &b->envs[env];
  int32_t* ids[2] = {ids_pred, ids_prey};
  int32_t* xy[2] = {xy_pred, xy_prey};
  double* en[2] = {energy_pred, energy_prey};
  if (b->cfg.variant == PPG_VARIANT_STAG) { /* agent_positions insertion order = self.agents order (flat ids are not monotonic) */
    n_live[0] = n_live[1] = 0;
    for (int i = 0; i < v->n_agents; ++i) {
      const int s = KEY_S(v->agents[i]), id = KEY_ID(v->agents[i]);
      if (!v->present[s][id]) continue;
      const int n = n_live[s]++;
      if (ids[s]) ids[s][n] = id;
      if (xy[s]) { xy[s][2 * n] = v->x[s][id]; xy[s][2 * n + 1] = v->y[s][id]; }
      if (en[s]) en[s][n] = v->energy[s][id];
    }
  } else
  for (int s = 0; s < 2; ++s) {
    int n = 0;
    for (int id = 0; id < v->next_idx[s]; ++id) /* agent_positions insertion order = id order */
      if (v->present[s][id]) {
        if (ids[s]) ids[s][n] = id;
        if (xy[s]) { xy[s][2 * n] = v->x[s][id]; xy[s][2 * n + 1] = v->y[s][id]; }
        if (en[s]) en[s][n] = v->energy[s][id];
        ++n;
      }
    n_live[s] = n;
  }
  for (int g = 0; g < b->cfg.n_grass; ++g) {
    if (xy_grass) { xy_grass[2 * g] = v->gx[g]; xy_grass[2 * g + 1] = v->gy[g]; }
    if (energy_grass) energy_grass[g] = v->ge[g];
  }
  return PPG_OK;
}

int ppgo_read_grid(ppgo_batch* b, int32_t env, double* grid_out) {
  if (env < 0 || env >= b->n_envs) return PPG_ERR_INVALID;
  env_t* v = &b->envs[env];
  if (b->cfg.variant != PPG_VARIANT_BASE) { eco_read_grid(v, grid_out); return PPG_OK; }
  memcpy(grid_out, v->grid, sizeof(double) * (size_t)v->C * v->G * v->G);
  return PPG_OK;
}

/* ---- ECO extras ---- */
int ppgo_env_reset_eco(ppgo_batch* b, int32_t env, const int32_t* cells, const double* founder_speed) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_ECO) return PPG_ERR_INVALID;
  for (int e = 0; e < b->n_envs; ++e) if (e != env) b->envs[e].n_rows = 0;
  b->envs[env].episode += 1;
  b->envs[env].trait_draws = 0;
  eco_env_reset_explicit(&b->envs[env], cells, founder_speed);
  export_rows(b);
  return PPG_OK;
}

/* trait variants: reset one env with an explicit number of founders (MR:189-192), cells and founder traits */
int ppgo_env_reset_trait(ppgo_batch* b, int32_t env, int32_t n_pred, int32_t n_prey, const int32_t* cells, const double* founder_trait) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_ECO || b->cfg.trait_mode == PPG_TRAIT_SPEED) return PPG_ERR_INVALID;
  if (n_pred < 0 || n_pred > b->cfg.n_initial[0] || n_prey < 0 || n_prey > b->cfg.n_initial[1]) return PPG_ERR_INVALID;
  for (int e = 0; e < b->n_envs; ++e) if (e != env) b->envs[e].n_rows = 0;
  b->envs[env].episode += 1;
  b->envs[env].trait_draws = 0;
  b->envs[env].n_found[0] = n_pred; b->envs[env].n_found[1] = n_prey;
  eco_env_reset_explicit(&b->envs[env], cells, founder_trait);
  export_rows(b);
  return PPG_OK;
}

int ppgo_read_env_eco(ppgo_batch* b, int32_t env, int32_t* age_pred, double* speed_pred, int32_t* age_prey,
                      double* speed_prey, uint8_t* dead_prey, int32_t* active_num) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_ECO) return PPG_ERR_INVALID;
  env_t* v = &b->envs[env];
  int32_t* age[2] = {age_pred, age_prey};
  double* sp[2] = {speed_pred, speed_prey};
  for (int s = 0; s < 2; ++s) {
    int n = 0;
    for (int id = 0; id < v->next_idx[s]; ++id)
      if (v->present[s][id]) {
        if (age[s]) age[s][n] = v->age[s][id];
        if (sp[s]) sp[s][n] = v->speed[s][id];
        if (s == 1 && dead_prey) dead_prey[n] = v->dead[id];
        ++n;
      }
  }
  if (active_num) { active_num[0] = v->active[0]; active_num[1] = v->active[1]; }
  return PPG_OK;
}

/* per-episode totals of one ECO env, the layout of ppg_read_episode_eco (include/ppg.h; ECO:1613-1661) */
int ppgo_read_episode_eco(ppgo_batch* b, int32_t env, double* sums, int32_t* spawned) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_ECO) return PPG_ERR_INVALID;
  const env_t* v = &b->envs[env];
  if (sums) for (int k = 0; k < 4; ++k) sums[k] = v->ep_sums[k];
  if (spawned) { spawned[0] = v->ep_spawned[0]; spawned[1] = v->ep_spawned[1]; }
  return PPG_OK;
}

/* the trait variants' event counters of one env, the layout of ppg_read_episode_events_eco (include/ppg.h; MR:1347-1350, COOP:1365-1368) */
int ppgo_read_episode_events_eco(ppgo_batch* b, int32_t env, double* events) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_ECO || !events) return PPG_ERR_INVALID;
  for (int k = 0; k < 6; ++k) events[k] = b->envs[env].ep_events[k];
  return PPG_OK;
}

/* CAD: agent_move_accumulator (CAD:183-186) of one env, in the order of ppgo_read_env_eco */
int ppgo_read_env_acc(ppgo_batch* b, int32_t env, double* acc_pred, double* acc_prey) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_ECO) return PPG_ERR_INVALID;
  env_t* v = &b->envs[env];
  double* out[2] = {acc_pred, acc_prey};
  for (int s = 0; s < 2; ++s) {
    int n = 0;
    for (int id = 0; id < v->next_idx[s]; ++id)
      if (v->present[s][id]) out[s][n++] = v->acc[s][id];
  }
  return PPG_OK;
}

/* ---- STAG extras ---- */
int ppgo_env_reset_stag(ppgo_batch* b, int32_t env, const int32_t* cells, const int32_t* facing, const double* trait_raw) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_STAG) return PPG_ERR_INVALID;
  for (int e = 0; e < b->n_envs; ++e) if (e != env) b->envs[e].n_rows = 0;
  b->envs[env].episode += 1;
  b->envs[env].trait_draws = 0;
  stag_env_reset_explicit(&b->envs[env], cells, facing, trait_raw);
  export_rows(b);
  return PPG_OK;
}

int ppgo_read_env_stag(ppgo_batch* b, int32_t env, int32_t* age_pred, int32_t* facing_pred, double* trait_pred,
                       int32_t* age_prey, int64_t* capture, double* capture_real) {
  if (env < 0 || env >= b->n_envs || b->cfg.variant != PPG_VARIANT_STAG) return PPG_ERR_INVALID;
  env_t* v = &b->envs[env];
  int n[2] = {0, 0};
  for (int i = 0; i < v->n_agents; ++i) {
    const int s = KEY_S(v->agents[i]), id = KEY_ID(v->agents[i]);
    if (!v->present[s][id]) continue;
    const int k = n[s]++;
    if (s == 0) {
      if (age_pred) age_pred[k] = v->age[0][id];
      if (facing_pred) facing_pred[k] = v->facing[id];
      if (trait_pred) trait_pred[k] = v->trait[id];
    } else if (age_prey) age_prey[k] = v->age[1][id];
  }
  if (capture) memcpy(capture, v->capture, sizeof v->capture);
  if (capture_real) memcpy(capture_real, v->capture_real, sizeof v->capture_real);
  return PPG_OK;
}

/* ---- literal single-env interface (golden tests) ---- */
int ppgo_env_reset_cells(ppgo_batch* b, int32_t env, const int32_t* cells) {
  if (env < 0 || env >= b->n_envs) return PPG_ERR_INVALID;
  for (int e = 0; e < b->n_envs; ++e) if (e != env) b->envs[e].n_rows = 0;
  b->envs[env].episode += 1;
  env_reset_cells(&b->envs[env], cells);
  export_rows(b);
  return PPG_OK;
}

int ppgo_env_step_ordered(ppgo_batch* b, int32_t env, int32_t n, const int32_t* species,
                          const int32_t* ids, const int32_t* actions) {
  if (env < 0 || env >= b->n_envs) return PPG_ERR_INVALID;
  for (int e = 0; e < b->n_envs; ++e) if (e != env) b->envs[e].n_rows = 0;
  int trunc_call = b->envs[env].current_step >= b->cfg.max_steps;
  int rc = env_step(&b->envs[env], n, species, ids, actions, 0);
  if (rc) return rc;
  export_rows(b);
  if (!trunc_call) env_sort_agents(&b->envs[env]);
  return PPG_OK;
}

int ppgo_env_agents(ppgo_batch* b, int32_t env, int32_t cap, int32_t* species, int32_t* ids) {
  if (env < 0 || env >= b->n_envs) return -1;
  env_t* v = &b->envs[env];
  for (int i = 0; i < v->n_agents && i < cap; ++i) { species[i] = KEY_S(v->agents[i]); ids[i] = KEY_ID(v->agents[i]); }
  return v->n_agents;
}
