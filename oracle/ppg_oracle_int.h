/*
 * ppg_oracle_int.h — internal types shared by the oracle's translation units (TEST INFRASTRUCTURE ONLY).
 */
#ifndef PPG_ORACLE_INT_H_
#define PPG_ORACLE_INT_H_

#include "ppg_oracle.h"

#define KEY(s, id) (((int32_t)(s) << 16) | (int32_t)(id))
#define KEY_S(k) ((k) >> 16)
#define KEY_ID(k) ((k)&0xFFFF)

/* ------------------------------------------------------------------------------------------ */
/* one environment instance                                                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct env_t {
  const ppg_config* c;
  int G, C;
  int env_index;
  /* dicts keyed by agent id: agent_positions / agent_energies (BASE:111-117) */
  uint8_t* present[2];
  int16_t* x[2];
  int16_t* y[2];
  double* energy[2];
  int32_t* parent[2]; /* KICK:86-91 agent_parent, -1 = none */
  /* self.agents (BASE:73), self._pending_removal (BASE:65) */
  int32_t* agents;
  int n_agents;
  int32_t* pending;
  int n_pending;
  int next_idx[2]; /* BASE:66-67 */
  int cur_num[2];  /* BASE:210-211 */
  int current_step;
  double* grid; /* [C][G][G] BASE:123-124 */
  int16_t* gx;
  int16_t* gy;
  double* ge; /* grass_positions / grass_energies */
  /* per-call dicts: observations / rewards / terminations, keyed by list index of self.agents */
  double* obs;     /* [cap_rows][max_row_elems] */
  double* rew;     /* rewards[agent] */
  int8_t* has_rew;
  int8_t* term;    /* -1 missing, 0 False, 1 True */
  int8_t* trunc;
  int8_t* has_obs;
  uint8_t* ate;    /* agents_just_ate */
  uint8_t* repro;  /* had an offspring this step (PPG_ROW_REPRODUCED) */
  uint8_t* newborn;
  double* e_before; /* ADD:256 energy_before */
  double* bonus;    /* ADD:261 reproduction_bonus */
  int32_t* list_index[2]; /* id -> index in self.agents during the call */
  int n_rows;      /* rows produced by the last call (= len(self.agents) at output time) */
  int cap_rows;
  int row_elems[2];
  int max_row_elems;
  uint8_t all_term, all_trunc; /* "__all__" */
  uint8_t env_flags;
  /* lockstep layer */
  int needs_reset, idle;
  uint8_t status;
  uint64_t seed_key;
  uint32_t episode, spawn_draws;
  const int32_t* tape_cells;
  int64_t tape_pos, tape_end;
  int64_t stats[PPG_N_STATS];
  /* lexicographic rank of str(id): Python sorts agent-id strings (BASE:468) */
  const int32_t* lexrank[2];
  /* ---- ECO (ppg_oracle_eco.c) ---- */
  int32_t* row_key;   /* ECO: agent of output row i (BASE: row i is self.agents[i]); NULL for BASE */
  uint8_t* carcass;   /* per row: in dead_prey after the step */
  uint8_t* born_obs;  /* per row: observation captured at birth (ECO:1179) is still the one in self.observations */
  uint8_t* frozen;    /* per row, CAD: the action mask captured with the observation allows only "stay" (CAD:746-753) */
  float* gridf;       /* ECO grid_world_state is float32 [C][G][G] (ECO:222-224) */
  int32_t* age[2];    /* agent_ages (ECO:153) */
  double* speed[2];   /* agent_genomes[..].speed (ECO:177); < 0 = no genome */
  uint8_t* termd[2];  /* self.terminations.get(agent) of the running step */
  uint8_t* dead;      /* dead_prey membership by prey id (ECO:193) */
  int active[2];      /* active_num_predators / active_num_prey (ECO:212-213) */
  const double* tape_reals;
  int64_t real_pos, real_end;
  uint32_t trait_draws;
  /* ---- trait variants of ECO (MR / INV / COOP, ppg_oracle_eco.c) ---- */
  int n_found[2];     /* founders of the running episode (MR:189-192) */
  int32_t* sat_until; /* agent_satiation_until by predator id (MR:134,756-757) */
  double* acc[2];     /* CAD: agent_move_accumulator by id (CAD:183-186) */
  /* per-episode totals behind _build_episode_training_metrics (ECO:1613-1661): distance moved and locomotion energy of all
   * agents of a species (the sums of record["distance_traveled"] / record["movement_energy_spent"], ECO:659-660), births */
  double ep_sums[4];
  /* trait variants' event counters in the layout of ppg_read_episode_events_eco (include/ppg.h): births blocked by the id pool
   * [2] (MR:857,943), by the density cap (MR:852), catches blocked by satiation (MR:739), energy donated [2] (COOP:585-586) */
  double ep_events[6];
  int32_t ep_spawned[2];
  /* ECO lineage_tracker by id (ECO:1422-1470): parent (-1: founder), live_descendants, prev_live_descendants, is_alive_descendant */
  int32_t* lin_parent[2];
  int32_t* lin_live[2];
  int32_t* lin_prev[2];
  uint8_t* lin_alive[2];
  /* ---- STAG (ppg_oracle_stag.c) ---- */
  uint8_t* wallmap;   /* [G*G] 1 at wall cells (STAG:2107-2127) */
  int8_t* facing;     /* predator_facing as an index into _predator_facing_options (STAG:197-206) */
  double* trait;      /* predator_cooperation_trait (STAG:230) */
  uint8_t* join;      /* predator_join_intent of the running step: 0 / 1, 2 = no entry (defaults to True, STAG:1163) */
  int next_idx_t[2][2]; /* head of the per-type id deques (STAG:2259-2291) */
  int64_t capture[12];  /* team-capture counters in the order of ppg_read_env_stag (include/ppg.h) */
  double capture_real[3];
  uint32_t facing_draws, capture_draws;
} env_t;

struct ppgo_batch {
  ppg_config cfg;
  int n_envs;
  env_t* envs;
  int32_t* lexrank[2];
  /* tape copy */
  int32_t* tape_cells;
  int64_t* tape_off;
  double* tape_reals;
  int64_t* tape_real_off;
  int has_tape;
  /* flat outputs */
  ppgo_buffers out;
  int64_t cap[2];
  int32_t n_rows[4];
  uint64_t calls;
  int n_threads;
  /* row -> (list position) bookkeeping of the previous output, for action lookup */
  int32_t* prev_row[2]; /* [env][id] flattened lazily: per env arrays */
};


/* ---- ECO (ppg_oracle_eco.c) ---- */
void eco_env_alloc(env_t* e);
void eco_env_free(env_t* e);
void eco_ensure_rows(env_t* e, int need);
void eco_env_reset_auto(env_t* e);
void eco_env_reset_explicit(env_t* e, const int32_t* cells, const double* founder_speed);
void eco_founder_counts(env_t* e, int from_tape);
int eco_env_step(env_t* e, int n_act, const int32_t* a_s, const int32_t* a_id, const int32_t* a_val);
void eco_read_grid(env_t* e, double* out);


/* ---- STAG (ppg_oracle_stag.c) ---- */
void stag_env_alloc(env_t* e);
void stag_env_free(env_t* e);
void stag_env_reset_auto(env_t* e);
void stag_env_reset_explicit(env_t* e, const int32_t* cells, const int32_t* facing, const double* trait_raw);
int stag_env_step(env_t* e, int n_act, const int32_t* a_s, const int32_t* a_id, const int32_t* a_val);

#endif
