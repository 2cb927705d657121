/*
 * ppg_oracle.h — CPU oracle of the PredPreyGrass environment step.  TEST INFRASTRUCTURE ONLY.
 *
 * A sequential, literal restatement in plain C of the reference's Python `reset()/step()`
 * (one interpreter per env instance, persistent float64 grid exactly as the reference keeps it),
 * plus a thin lockstep layer that lays the per-env results out in the row format of
 * include/ppg.h so the CUDA path can be compared array against array.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (predpreygrass_b200/) never links,
 * imports or calls it.
 *
 * Parity pin: tests/golden/*.npz hold trajectories recorded from the UNMODIFIED reference classes
 * (imported from /root/reference under the tests/golden/_shim stubs by tests/golden/make_golden.py);
 * tests/test_oracle_golden.py replays them through this oracle bit for bit.
 */
#ifndef PPG_ORACLE_H_
#define PPG_ORACLE_H_

#include "../include/ppg.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppgo_batch ppgo_batch;

/* host-side mirror of ppg_buffers: same meaning, host pointers, obs additionally in float64 */
typedef struct ppgo_buffers {
  ppg_buffers f;  /* float32 obs etc. (host pointers) */
  double* obs64[2];    /* the reference's own dtype for BASE (float64, BASE:518-521) */
  double* reward64[2]; /* rewards before the float32 cast */
} ppgo_buffers;

ppgo_batch* ppgo_create(const ppg_config* cfg, int32_t n_envs);
void ppgo_destroy(ppgo_batch* b);
int ppgo_load_tape(ppgo_batch* b, const ppg_tape* tape);
int ppgo_reset(ppgo_batch* b, const uint64_t* seeds, const uint8_t* mask);
/* actions indexed by the rows of the previous output, like ppg_step */
int ppgo_step(ppgo_batch* b, const int32_t* actions_pred, const int32_t* actions_prey);
int ppgo_random_actions(ppgo_batch* b, uint64_t seed, int32_t* actions_pred, int32_t* actions_prey);
int ppgo_get_buffers(ppgo_batch* b, ppgo_buffers* out);
int ppgo_stats(ppgo_batch* b, int64_t* out);
int ppgo_read_env(ppgo_batch* b, int32_t env, int32_t* n_live, int32_t* ids_pred, int32_t* xy_pred,
                  double* energy_pred, int32_t* ids_prey, int32_t* xy_prey, double* energy_prey,
                  int32_t* xy_grass, double* energy_grass);
/* persistent grid of one env, float64 [C][G][G] (BASE:124) */
int ppgo_read_grid(ppgo_batch* b, int32_t env, double* grid_out);
/* number of worker threads for ppgo_step (envs are independent); default 1 */
void ppgo_set_threads(ppgo_batch* b, int32_t n_threads);

/*
 * Literal single-env interface used by the golden-vector tests: the caller supplies the action
 * dict in ITS OWN iteration order (species[i], id[i], action[i]), exactly like
 * `env.step(action_dict)`; includes the BASE extra truncation call (BASE:228-238).
 * Results are read back through ppgo_get_buffers (env 0 rows, reference dict order).
 * Returns 0, or -1 if an action key is not a live agent (reference: KeyError, BASE:246).
 */
int ppgo_env_step_ordered(ppgo_batch* b, int32_t env, int32_t n, const int32_t* species,
                          const int32_t* ids, const int32_t* actions);
/* reset one env from explicit cells (predators, prey, grass order; BASE:185-187) */
int ppgo_env_reset_cells(ppgo_batch* b, int32_t env, const int32_t* cells);
/* `self.agents` of one env (BASE:73,468): returns count, fills species/ids (capacity cap) */
int ppgo_env_agents(ppgo_batch* b, int32_t env, int32_t cap, int32_t* species, int32_t* ids);

/* ---- ECO (ppg_oracle_eco.c) ---- */
/* reset one env from explicit cells (predators, prey, grass; ECO:1752-1762) and founder speeds (genome.py:42-46) */
int ppgo_env_reset_eco(ppgo_batch* b, int32_t env, const int32_t* cells, const double* founder_speed);
/* trait variants (MR / INV / COOP): reset one env with an explicit founder count per species, cells and founder traits */
int ppgo_env_reset_trait(ppgo_batch* b, int32_t env, int32_t n_pred, int32_t n_prey, const int32_t* cells, const double* founder_trait);
/* CAD: the move accumulators of one env, in the order of ppgo_read_env_eco */
/* per-episode totals of one ECO env in the layout of ppg_read_episode_eco (include/ppg.h): sums[4] = distance pred / prey, locomotion
 * energy pred / prey; spawned[2] = births so far (ECO:1613-1661) */
int ppgo_read_episode_eco(ppgo_batch* b, int32_t env, double* sums, int32_t* spawned);
/* the trait variants' event counters in the layout of ppg_read_episode_events_eco (include/ppg.h): events[6] */
int ppgo_read_episode_events_eco(ppgo_batch* b, int32_t env, double* events);
int ppgo_read_env_acc(ppgo_batch* b, int32_t env, double* acc_pred, double* acc_prey);
/* agent_ages / genome speeds / dead_prey / active_num_* of one env, in the list order of ppgo_read_env */
int ppgo_read_env_eco(ppgo_batch* b, int32_t env, int32_t* age_pred, double* speed_pred, int32_t* age_prey,
                      double* speed_prey, uint8_t* dead_prey, int32_t* active_num);

/* ---- STAG (ppg_oracle_stag.c) ---- */
/* reset one env from explicit cells, founder facing indices and raw (unclipped) founder traits (STAG:2140,2168-2169) */
int ppgo_env_reset_stag(ppgo_batch* b, int32_t env, const int32_t* cells, const int32_t* facing, const double* trait_raw);
/* same layout as ppg_read_env_stag (include/ppg.h) */
int ppgo_read_env_stag(ppgo_batch* b, int32_t env, int32_t* age_pred, int32_t* facing_pred, double* trait_pred,
                       int32_t* age_prey, int64_t* capture, double* capture_real);

#ifdef __cplusplus
}
#endif
#endif
