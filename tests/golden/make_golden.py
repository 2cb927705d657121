#!/usr/bin/env python
"""Record golden trajectories from the UNMODIFIED reference env classes.

Runs in the build container only (needs /root/reference); the GPU box never executes this.
Usage:  python tests/golden/make_golden.py            (re-creates tests/golden/base_*.npz)

For every case the unmodified reference class (imported from /root/reference under the stub
`ray`/`gymnasium` packages in tests/golden/_shim) is stepped with seeded uniform-random actions and
the following is recorded, all in the reference's own iteration orders:
  * the RNG tape: initial cells after reset() (predators, prey, grass) and every spawn-fallback
    cell (`np.random.randint` draw, BASE:764),
  * the action dict of every step as (species, id, action) in the order it was passed,
  * the returned dicts: observation-dict key order, rewards, terminations, truncations, "__all__",
    and a sha1 over the float64 observation bytes (dict order),
  * the state after the step: agent_positions/agent_energies (insertion order), grass energies,
    sha1 of grid_world_state, self.agents.
The oracle (oracle/ppg_oracle.c) must reproduce all of it bit for bit (tests/test_oracle_golden.py).
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

NE = "predpreygrass.non_evolutionary."
RS = NE + "project_reward_shaping."
MODULES = {
    "sparse": NE + "base_environment",
    "dense": RS + "base_environment_dense_rewards",
    "additive": RS + "base_environment_dense_rewards_additive",
    "kickback": RS + "base_environment_sparse_rewards_plus_kickback",
    "eating": RS + "base_environment_sparse_rewards_plus_eating",
    "seasonal": NE + "base_environment_seasonal",
}

# small, crowded world: frequent births, blocked moves, co-located agents, spawn fallback draws
CROWDED = dict(grid_size=6, initial_num_grass=14, n_initial_active_predator=5, n_initial_active_prey=12,
               predator_creation_energy_threshold=6.0, prey_creation_energy_threshold=3.5,
               initial_energy_predator=3.0, initial_energy_prey=2.0, energy_gain_per_step_grass=0.5,
               energy_loss_per_step_prey=0.02, energy_loss_per_step_predator=0.1,
               predator_obs_range=5, prey_obs_range=7, n_possible_predators=300, n_possible_prey=300)

CASES = [
    # name, variant, overrides, seed, action order ("dict" = previous obs-dict order, "shuffle", "agents"), max_calls
    ("base_default_s1", "sparse", {}, 1, "dict", 400),
    ("base_default_s3", "sparse", {}, 3, "dict", 400),
    ("base_default_s7_shuffle", "sparse", {}, 7, "shuffle", 300),
    ("base_trunc_s2", "sparse", {"max_steps": 40}, 2, "dict", 60),
    ("base_crowded_s1", "sparse", CROWDED, 1, "dict", 250),
    ("base_crowded_s2_shuffle", "sparse", CROWDED, 2, "shuffle", 250),
    ("base_founders12_s4", "sparse", {"n_initial_active_predator": 12, "n_initial_active_prey": 14}, 4, "dict", 200),
    ("eating_default_s5", "eating", {}, 5, "dict", 250),
    ("dense_default_s1", "dense", {}, 1, "dict", 250),
    ("dense_crowded_s3", "dense", CROWDED, 3, "dict", 200),
    ("additive_default_s1", "additive", {}, 1, "dict", 300),
    ("additive_crowded_s2", "additive", CROWDED, 2, "dict", 200),
    ("kickback_crowded_s1", "kickback", CROWDED, 1, "dict", 200),
    ("kickback_default_s6", "kickback", {}, 6, "dict", 300),
    # every reward constant non-zero (BASE:288,322,328,341,365,375): step rewards, catch / eat rewards, caught penalty
    ("base_rewards_s8", "sparse", dict(CROWDED, reward_predator_step=0.125, reward_prey_step=0.25, reward_predator_catch_prey=3.0,
                                       reward_prey_eat_grass=1.5, penalty_prey_caught=-2.0, reproduction_reward_predator=7.0,
                                       reproduction_reward_prey=4.0, grid_size=8, initial_num_grass=24), 8, "dict", 200),
    # other window shapes than the default 7 / 9 (generic row writer on the device), odd grid size
    ("base_wide_s9", "sparse", dict(grid_size=15, predator_obs_range=9, prey_obs_range=9, initial_num_grass=40,
                                    n_possible_predators=200, n_possible_prey=300), 9, "shuffle", 200),
    ("base_narrow_s10", "sparse", dict(grid_size=11, predator_obs_range=3, prey_obs_range=5, initial_num_grass=30,
                                       n_possible_predators=200, n_possible_prey=300, energy_gain_per_step_grass=0.2), 10, "dict", 200),
    # base_environment_seasonal: square-wave multiplier on the grass regrowth (season_length_steps, high / low multiplier)
    ("seasonal_default_s1", "seasonal", {}, 1, "dict", 300),
    ("seasonal_crowded_s2_shuffle", "seasonal", dict(CROWDED, season_length_steps=7, season_high_multiplier=2.0, season_low_multiplier=0.25),
     2, "shuffle", 200),
    ("seasonal_trunc_s3", "seasonal", {"max_steps": 25, "season_length_steps": 10}, 3, "dict", 60),
    ("seasonal_default_s5", "seasonal", {"season_length_steps": 25, "season_high_multiplier": 2.5, "season_low_multiplier": 0.0}, 5, "dict", 400),
]


def split(agent):
    kind, idx = agent.rsplit("_", 1)
    return (0 if kind == "predator" else 1), int(idx)


def sha(arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def record(name, variant, overrides, seed, order, max_calls):
    mod = importlib.import_module(MODULES[variant] + ".predpreygrass_rllib_env")
    cfgmod = importlib.import_module(MODULES[variant] + ".config_env")
    cfg = dict(cfgmod.config_env)
    cfg.update(overrides)
    env = mod.PredPreyGrass(cfg)
    G = env.grid_size

    fallback = []
    real_randint = np.random.randint
    orig_find = env._find_available_spawn_position

    def find_wrapped(ref_pos, occupied):
        drew = []

        def ri(*a, **k):
            drew.append(1)
            return real_randint(*a, **k)

        np.random.randint = ri
        try:
            pos = orig_find(ref_pos, occupied)
        finally:
            np.random.randint = real_randint
        if drew:
            fallback.append(int(pos[0]) * G + int(pos[1]))
        return pos

    env._find_available_spawn_position = find_wrapped  # instrumentation of the instance, source untouched
    np.random.seed(seed + 1000)  # BASE:764 uses the global numpy RNG

    obs, _ = env.reset(seed=seed)
    init_cells = [int(env.agent_positions[a][0]) * G + int(env.agent_positions[a][1]) for a in env.agents]
    init_cells += [int(p[0]) * G + int(p[1]) for p in env.grass_positions.values()]
    reset_keys = [split(a) for a in obs]
    reset_sha = sha([obs[a] for a in obs])

    arng = np.random.default_rng(seed * 7919 + 13)
    rec = {k: [] for k in ("act_s", "act_id", "act_v", "row_s", "row_id", "row_rew", "row_term", "row_trunc",
                           "st_s", "st_id", "st_x", "st_y", "st_e", "ag_s", "ag_id")}
    offs = {k: [0] for k in ("act", "row", "st", "ag")}
    obs_sha, grid_sha, grass_e, all_term, all_trunc, steps = [], [], [], [], [], []
    done = False
    calls = 0
    while not done and calls < max_calls:
        if order == "agents":
            keys = [a for a in env.agents if a in env.agent_positions]
        else:
            keys = [a for a in obs if a in env.agent_positions]
        if order == "shuffle":
            keys = [keys[i] for i in arng.permutation(len(keys))]
        acts = {a: int(arng.integers(9)) for a in keys}
        obs, rew, term, trunc, _ = env.step(acts)
        calls += 1
        assert list(obs) == list(rew) == [k for k in term if k != "__all__"] == [k for k in trunc if k != "__all__"], name
        for a, v in acts.items():
            s, i = split(a)
            rec["act_s"].append(s); rec["act_id"].append(i); rec["act_v"].append(v)
        offs["act"].append(len(rec["act_s"]))
        for a in obs:
            s, i = split(a)
            rec["row_s"].append(s); rec["row_id"].append(i); rec["row_rew"].append(float(rew[a]))
            rec["row_term"].append(int(bool(term[a]))); rec["row_trunc"].append(int(bool(trunc[a])))
        offs["row"].append(len(rec["row_s"]))
        for a, p in env.agent_positions.items():
            s, i = split(a)
            rec["st_s"].append(s); rec["st_id"].append(i); rec["st_x"].append(int(p[0])); rec["st_y"].append(int(p[1]))
            rec["st_e"].append(float(env.agent_energies[a]))
        offs["st"].append(len(rec["st_s"]))
        for a in env.agents:
            s, i = split(a)
            rec["ag_s"].append(s); rec["ag_id"].append(i)
        offs["ag"].append(len(rec["ag_s"]))
        obs_sha.append(sha([obs[a] for a in obs]))
        grid_sha.append(sha([env.grid_world_state]))
        grass_e.append([float(env.grass_energies[g]) for g in env.grass_agents])
        all_term.append(int(bool(term["__all__"]))); all_trunc.append(int(bool(trunc["__all__"])))
        steps.append(env.current_step)
        done = term["__all__"] or trunc["__all__"]

    out = dict(
        cfg_json=np.array(json.dumps(dict(cfg, variant=variant))), seed=np.int64(seed), order=np.array(order),
        init_cells=np.array(init_cells, np.int32), fallback_cells=np.array(fallback, np.int32),
        reset_row_s=np.array([k[0] for k in reset_keys], np.int8), reset_row_id=np.array([k[1] for k in reset_keys], np.int32),
        reset_sha=reset_sha,
        act_off=np.array(offs["act"], np.int64), row_off=np.array(offs["row"], np.int64),
        st_off=np.array(offs["st"], np.int64), ag_off=np.array(offs["ag"], np.int64),
        act_s=np.array(rec["act_s"], np.int8), act_id=np.array(rec["act_id"], np.int32), act_v=np.array(rec["act_v"], np.int8),
        row_s=np.array(rec["row_s"], np.int8), row_id=np.array(rec["row_id"], np.int32),
        row_rew=np.array(rec["row_rew"], np.float64), row_term=np.array(rec["row_term"], np.int8),
        row_trunc=np.array(rec["row_trunc"], np.int8),
        st_s=np.array(rec["st_s"], np.int8), st_id=np.array(rec["st_id"], np.int32), st_x=np.array(rec["st_x"], np.int16),
        st_y=np.array(rec["st_y"], np.int16), st_e=np.array(rec["st_e"], np.float64),
        ag_s=np.array(rec["ag_s"], np.int8), ag_id=np.array(rec["ag_id"], np.int32),
        obs_sha=np.array(obs_sha, np.uint8).reshape(-1, 20), grid_sha=np.array(grid_sha, np.uint8).reshape(-1, 20),
        grass_e=np.array(grass_e, np.float64), all_term=np.array(all_term, np.int8), all_trunc=np.array(all_trunc, np.int8),
        steps=np.array(steps, np.int32),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    births = int(env._next_predator_idx - env.n_initial_active_predator + env._next_prey_idx - env.n_initial_active_prey)
    print(f"{name:28s} calls={calls:4d} births={births:4d} fallback={len(fallback):3d} "
          f"end={'term' if all_term[-1] else ('trunc' if all_trunc[-1] else 'cut')} size={os.path.getsize(path)//1024}KB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        record(*case)
