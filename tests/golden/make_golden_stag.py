#!/usr/bin/env python
"""Record golden trajectories from the UNMODIFIED reference STAG env class
(predpreygrass/evolutionary/stag_hunt_forward_view_nature_nurture/predpreygrass_rllib_env.py).

Runs in the build container only (needs /root/reference); the GPU box never executes this.
Usage:  python tests/golden/make_golden_stag.py [case ...]     (re-creates tests/golden/stag_*.npz)

The class is imported under the stub `ray`/`gymnasium` packages of tests/golden/_shim and stepped with seeded
uniform-random actions (predators: `[move, join_hunt]`).  Agents are keyed (species, flat id) with
flat id = k for `type_1_*_k` and n_possible_type_1_<species> + k for `type_2_*_k` (the numbering of include/ppg.h).
Recorded per case:
  * the RNG tape.  ints: initial cells (STAG:2140 `rng.choice`), founders' facing indices (STAG:941,2168), then in
    consumption order the cell of every spawn-fallback draw (STAG:1039-1042) and every newborn predator's facing
    index (STAG:1559).  reals: founders' cooperation traits before clipping (STAG:1087), then per predator birth
    `u = rng.random()` and, iff u < rate, `delta = rng.normal(0, std)` (STAG:1095-1096), and per capture attempt
    `u = rng.random()` iff the success model draws (STAG:1146,1148) — `env.rng` is wrapped in a recording proxy
    on the instance, the source is untouched;
  * the action dict of every step as (species, id, move, join) in the order passed (STAG:805);
  * the returned dicts: observation-dict order of the live agents (= self.agents, STAG:551) followed by the ended
    agents sorted by key (the reference adds those from a Python set, STAG:597-612), rewards, terminations,
    truncations, float32 observations (sha1 over all + full arrays of sampled steps), "__all__";
  * the state after the step: agent_positions (insertion order), energies, ages, facing, trait, grass energies, sha1
    of the float32 grid, active_num_*, the team-capture counters and last success probability / effort ratio.
The oracle (oracle/ppg_oracle_stag.c) must reproduce all of it bit for bit (tests/test_oracle_golden_stag.py).
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PPG_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REF)

STAG = "predpreygrass.evolutionary.stag_hunt_forward_view_nature_nurture"

# small, dense world: captures, failed hunts, births, spawn fallback and blocked moves every few steps
CROWDED = dict(
    grid_size=8, initial_num_grass=24, predator_obs_range=7, prey_obs_range=5,
    n_initial_active_type_1_predator=8, n_initial_active_type_1_prey=6, n_initial_active_type_2_prey=14,
    n_possible_type_1_predators=300, n_possible_type_1_prey=200, n_possible_type_2_prey=400,
    energy_treshold_creation_predator=6.0, energy_treshold_creation_prey={"type_1_prey": 12.0, "type_2_prey": 2.4},
    energy_gain_per_step_grass=0.3, initial_energy_predator=3.0,
)
RICH = dict(energy_gain_per_step_grass=0.2, energy_treshold_creation_predator=7.0,
            energy_treshold_creation_prey={"type_1_prey": 14.0, "type_2_prey": 2.4})

CASES = [
    # name, overrides, seed, action order ("list" = self.agents order, "shuffle"), max_calls
    ("stag_default_s1", {}, 1, "list", 300),
    ("stag_default_s2_shuffle", {}, 2, "shuffle", 300),
    ("stag_rich_s3", RICH, 3, "list", 250),
    ("stag_crowded_s1", CROWDED, 1, "list", 200),
    ("stag_crowded_s2_shuffle", dict(CROWDED, coop_trait_mutation_rate=0.5, coop_trait_mutation_std=0.2), 2, "shuffle", 200),
    ("stag_joincost_s3", dict(CROWDED, team_capture_join_cost=0.9, energy_loss_per_step_predator=0.02), 3, "list", 200),
    ("stag_propsplit_s4", dict(CROWDED, team_capture_equal_split=False, team_capture_scavenger_fraction=0.5,
                               team_capture_margin=0.5), 4, "shuffle", 200),
    ("stag_determ_s5", dict(CROWDED, team_capture_success_model="deterministic"), 5, "list", 200),
    ("stag_probab_s6", dict(CROWDED, team_capture_success_model="probabilistic", team_capture_min_success_prob=0.2,
                            team_capture_base_success_p0=0.3), 6, "list", 200),
    ("stag_notrait_s7", dict(CROWDED, coop_trait_enabled=False), 7, "list", 200),
    ("stag_type2pred_s8", dict(CROWDED, n_initial_active_type_2_predator=4, n_possible_type_2_predators=100, type_2_action_range=5,
                               reproduction_reward_predator={"type_1_predator": 10.0, "type_2_predator": 3.0}), 8, "shuffle", 200),
    ("stag_penalty_s9", dict(CROWDED, death_penalty_predator=-2.0, death_penalty_type_1_prey=-3.0, death_penalty_type_2_prey=-1.5,
                             strict_rllib_output=False), 9, "list", 200),
    ("stag_trunc_s2", dict(RICH, max_steps=30), 2, "list", 60),
    # other window shapes and an odd grid: predators 5x5 (forward shift 2), prey 9x9
    # walls and line of sight (STAG:860-925,977-994,1026-1037,2107-2160)
    ("stag_walls_s1", dict(CROWDED, grid_size=10, manual_wall_positions=[(2, 2), (2, 3), (2, 4), (5, 5), (5, 6), (6, 5), (8, 1), (0, 9), (12, 3)]),
     1, "list", 200),
    ("stag_walls_los_s2_shuffle", dict(CROWDED, grid_size=10, type_2_action_range=5, respect_los_for_movement=True,
                                       include_visibility_channel=True, mask_observation_with_visibility=True,
                                       manual_wall_positions=[(3, y) for y in range(2, 8)] + [(6, 1), (6, 2), (7, 7), (8, 7), (1, 8)]),
     2, "shuffle", 200),
    ("stag_walls_vis_s3", dict(RICH, include_visibility_channel=True, respect_los_for_movement=True, type_2_action_range=5,
                               manual_wall_positions=[(x, 15) for x in range(5, 25)] + [(10, y) for y in range(3, 12)]), 3, "list", 120),
    ("stag_windows_s10", dict(CROWDED, grid_size=11, predator_obs_range=5, prey_obs_range=9, initial_num_grass=30), 10, "shuffle", 200),
    ("stag_tiny_s3", dict(CROWDED, grid_size=5, initial_num_grass=8, n_initial_active_type_1_predator=5, n_initial_active_type_1_prey=3,
                          n_initial_active_type_2_prey=8, predator_obs_range=9, prey_obs_range=7), 3, "shuffle", 150),
]


def make_split(env):
    n1 = (env.n_possible_type_1_predators, env.n_possible_type_1_prey)

    def split(agent):
        parts = str(agent).split("_")  # type_{t}_{species}_{k}
        t, kind, k = int(parts[1]), parts[2], int(parts[3])
        s = 0 if kind == "predator" else 1
        return s, k + (n1[s] if t == 2 else 0)

    return split


def sha(arrs, dtype):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=dtype).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


class RecordingRng:
    """Delegates to the env's numpy Generator and logs the draws of the step path."""

    def __init__(self, rng, log):
        self._rng, self._log = rng, log

    def random(self, *a, **k):
        v = self._rng.random(*a, **k)
        self._log.append(("u", float(v)))
        return v

    def normal(self, *a, **k):
        v = self._rng.normal(*a, **k)
        self._log.append(("n", float(v)))
        return v

    def integers(self, *a, **k):
        v = self._rng.integers(*a, **k)
        self._log.append(("i", int(v)))
        return v

    def __getattr__(self, name):
        return getattr(self._rng, name)


FACINGS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]


def record(name, overrides, seed, order, max_calls):
    mod = importlib.import_module(STAG + ".predpreygrass_rllib_env")
    cfgmod = importlib.import_module(STAG + ".config.config_env_stag_hunt_forward_view")
    cfg = dict(cfgmod.config_env)
    cfg.update(overrides)
    env = mod.PredPreyGrass(cfg)
    split = make_split(env)
    G = env.grid_size

    # reset under a recording proxy: np.random.default_rng is created inside _init_reset_variables, so wrap the normal()
    # draws by re-deriving them: the founders' raw traits are re-drawn from a twin generator in the same order
    obs, _ = env.reset(seed=seed)
    twin = np.random.default_rng(seed)
    n_all = len(env.agents) + len(env.grass_agents)
    twin_cells = twin.choice([i for i in range(G * G) if (i // G, i % G) not in env.wall_positions], size=n_all, replace=False)  # STAG:2138-2140
    founders = list(env.agents)
    pred_founders = [a for a in founders if "predator" in a]
    founder_facing, founder_trait_raw = [], []
    for a in pred_founders:  # STAG:2168-2169
        founder_facing.append(int(twin.integers(8)))
        if env.coop_trait_enabled:
            founder_trait_raw.append(float(twin.normal(env.coop_trait_init_mean, env.coop_trait_init_std)))
    assert twin.bit_generator.state == env.rng.bit_generator.state, "twin generator out of step with the reference's reset"
    init_cells = [int(env.agent_positions[a][0]) * G + int(env.agent_positions[a][1]) for a in founders]
    init_cells += [int(p[0]) * G + int(p[1]) for p in env.grass_positions.values()]
    assert init_cells == [int(c) for c in twin_cells]
    for a, f, t in zip(pred_founders, founder_facing, founder_trait_raw or [None] * len(pred_founders)):
        assert env.predator_facing[a] == FACINGS[f]
        if t is not None:
            assert env.predator_cooperation_trait[a] == min(max(t, 0.0), 1.0)

    log = []
    env.rng = RecordingRng(env.rng, log)  # instrumentation of the instance, source untouched
    orig_find = env._find_available_spawn_position
    n_fallback = [0]

    def find_wrapped(ref_pos, occupied):
        n0 = len(log)
        pos = orig_find(ref_pos, occupied)
        drew = [e for e in log[n0:] if e[0] == "i"]
        if drew:  # replace the index into the sorted free list by the chosen cell
            assert len(log) == n0 + 1
            log[n0] = ("c", int(pos[0]) * G + int(pos[1]))
            n_fallback[0] += 1
        return pos

    env._find_available_spawn_position = find_wrapped

    def obs_rows(obs_dict, live_order):
        live = [a for a in live_order if a in obs_dict]
        ended = sorted((a for a in obs_dict if a not in set(live)), key=split)
        return live + ended

    reset_keys = [split(a) for a in obs]
    assert list(obs) == founders
    reset_obs = [obs[a] for a in obs]

    arng = np.random.default_rng(seed * 7919 + 13)
    names = ("act_s", "act_id", "act_move", "act_join", "row_s", "row_id", "row_rew", "row_term", "row_trunc", "st_s", "st_id", "st_x",
             "st_y", "st_e", "st_age", "st_face", "st_trait", "ag_s", "ag_id")
    rec = {k: [] for k in names}
    offs = {k: [0] for k in ("act", "row", "st", "ag")}
    obs_sha, grid_sha, grass_e, all_term, all_trunc, steps, active, counters, lastp = [], [], [], [], [], [], [], [], []
    full_obs = {}
    done = False
    calls = 0
    while not done and calls < max_calls:
        keys = list(env.agents)
        if order == "shuffle":
            keys = [keys[i] for i in arng.permutation(len(keys))]
        acts = {}
        for a in keys:
            n_moves = max(1, (env.type_1_act_range if "type_1" in a else env.type_2_act_range) ** 2)
            mv = int(arng.integers(n_moves))
            if "predator" in a:
                jn = int(arng.integers(2))
                acts[a] = [mv, jn] if calls % 3 else np.array([mv, jn])  # both accepted forms (STAG:777-781)
            else:
                jn = -1
                acts[a] = mv
            s, i = split(a)
            rec["act_s"].append(s); rec["act_id"].append(i); rec["act_move"].append(mv); rec["act_join"].append(jn)
        offs["act"].append(len(rec["act_s"]))
        live_before = list(env.agents)
        obs, rew, term, trunc, _ = env.step(acts)
        live_after = [a for a in env.agents if a in env.agent_positions]
        rows = obs_rows(obs, live_after)
        assert set(rows) == set(rew) == set(a for a in term if a != "__all__") == set(a for a in trunc if a != "__all__"), (name, calls)
        row_obs = ([], [])
        for a in rows:
            s, i = split(a)
            rec["row_s"].append(s); rec["row_id"].append(i); rec["row_rew"].append(float(rew[a]))
            rec["row_term"].append(int(bool(term[a]))); rec["row_trunc"].append(int(bool(trunc[a])))
            assert obs[a].dtype == np.float32
            row_obs[s].append(obs[a])
        offs["row"].append(len(rec["row_s"]))
        obs_sha.append(sha(row_obs[0] + row_obs[1], np.float32))
        if calls < 2 or calls % 50 == 0:
            full_obs[calls] = row_obs
        for a, p in env.agent_positions.items():
            s, i = split(a)
            rec["st_s"].append(s); rec["st_id"].append(i); rec["st_x"].append(int(p[0])); rec["st_y"].append(int(p[1]))
            rec["st_e"].append(float(env.agent_energies[a])); rec["st_age"].append(int(env.agent_ages[a]))
            f = env.predator_facing.get(a)
            rec["st_face"].append(FACINGS.index(tuple(f)) if f is not None else -1)
            rec["st_trait"].append(float(env.predator_cooperation_trait.get(a, -1.0)))
        offs["st"].append(len(rec["st_s"]))
        for a in live_after:
            s, i = split(a)
            rec["ag_s"].append(s); rec["ag_id"].append(i)
        offs["ag"].append(len(rec["ag_s"]))
        assert env.grid_world_state.dtype == np.float32
        grid_sha.append(sha([env.grid_world_state], np.float32))
        grass_e.append([float(env.grass_energies[g]) for g in env.grass_agents])
        all_term.append(int(bool(term["__all__"]))); all_trunc.append(int(bool(trunc["__all__"])))
        steps.append(env.current_step)
        active.append([int(env.active_num_predators), int(env.active_num_prey)])
        counters.append([env.team_capture_successes, env.team_capture_failures, env.team_capture_coop_successes,
                         env.team_capture_coop_failures, env.team_capture_mammoth_successes, env.team_capture_mammoth_failures,
                         env.team_capture_rabbit_successes, env.team_capture_rabbit_failures, env.team_capture_attempts,
                         env.team_capture_helper_total, env.spawned_predators, env.spawned_prey])
        lastp.append([float(env.team_capture_last_success_prob), float(env.team_capture_last_effort_ratio),
                      float(env.team_capture_success_prob_sum)])
        done = term["__all__"] or trunc["__all__"]
        calls += 1

    ints = [v for k, v in log if k in ("i", "c")]
    reals = [v for k, v in log if k in ("u", "n")]
    out = dict(
        cfg_json=np.array(json.dumps(dict(cfg, variant="stag"), default=lambda o: None)), seed=np.int64(seed), order=np.array(order),
        init_cells=np.array(init_cells, np.int32), founder_facing=np.array(founder_facing, np.int32),
        founder_trait_raw=np.array(founder_trait_raw, np.float64), step_ints=np.array(ints, np.int32),
        step_reals=np.array(reals, np.float64),
        reset_row_s=np.array([k[0] for k in reset_keys], np.int8), reset_row_id=np.array([k[1] for k in reset_keys], np.int32),
        reset_sha=sha(reset_obs, np.float32),
        act_off=np.array(offs["act"], np.int64), row_off=np.array(offs["row"], np.int64),
        st_off=np.array(offs["st"], np.int64), ag_off=np.array(offs["ag"], np.int64),
        act_s=np.array(rec["act_s"], np.int8), act_id=np.array(rec["act_id"], np.int32), act_move=np.array(rec["act_move"], np.int8),
        act_join=np.array(rec["act_join"], np.int8),
        row_s=np.array(rec["row_s"], np.int8), row_id=np.array(rec["row_id"], np.int32),
        row_rew=np.array(rec["row_rew"], np.float64), row_term=np.array(rec["row_term"], np.int8),
        row_trunc=np.array(rec["row_trunc"], np.int8),
        st_s=np.array(rec["st_s"], np.int8), st_id=np.array(rec["st_id"], np.int32), st_x=np.array(rec["st_x"], np.int16),
        st_y=np.array(rec["st_y"], np.int16), st_e=np.array(rec["st_e"], np.float64), st_age=np.array(rec["st_age"], np.int32),
        st_face=np.array(rec["st_face"], np.int8), st_trait=np.array(rec["st_trait"], np.float64),
        ag_s=np.array(rec["ag_s"], np.int8), ag_id=np.array(rec["ag_id"], np.int32),
        obs_sha=np.array(obs_sha, np.uint8).reshape(-1, 20), grid_sha=np.array(grid_sha, np.uint8).reshape(-1, 20),
        grass_e=np.array(grass_e, np.float64), all_term=np.array(all_term, np.int8), all_trunc=np.array(all_trunc, np.int8),
        steps=np.array(steps, np.int32), active=np.array(active, np.int32).reshape(-1, 2),
        counters=np.array(counters, np.int64).reshape(-1, 12), lastp=np.array(lastp, np.float64).reshape(-1, 3),
        full_obs_steps=np.array(sorted(full_obs), np.int32),
    )
    for t, rows in full_obs.items():
        for s in range(2):
            out[f"full_obs_{t}_{s}"] = np.stack(rows[s]) if rows[s] else np.zeros((0,), np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    c = counters[-1]
    print(f"{name:26s} calls={calls:4d} births={c[10]:3d}+{c[11]:3d} capt ok/fail={c[0]:3d}/{c[1]:3d} coop={c[2]:3d}/{c[3]:3d} "
          f"fallback={n_fallback[0]:3d} reals={len(reals):4d} "
          f"end={'term' if all_term[-1] else ('trunc' if all_trunc[-1] else 'cut')} size={os.path.getsize(path)//1024}KB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        record(*case)
