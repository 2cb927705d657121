#!/usr/bin/env python
"""Record golden trajectories from the UNMODIFIED reference ECO env class
(predpreygrass/evolutionary/eco_evolutionary/predpreygrass_rllib_env.py).

Runs in the build container only (needs /root/reference); the GPU box never executes this.
Usage:  python tests/golden/make_golden_eco.py [case ...]     (re-creates tests/golden/eco_*.npz)

The reference class is imported under the stub `ray`/`gymnasium` packages of tests/golden/_shim and
stepped with seeded uniform-random actions.  Recorded per case, in the reference's own orders:
  * the RNG tape:  founder speeds (`agent_genomes` after reset, genome.py:45), initial cells
    (ECO:1752 `rng.choice`), per birth `u = rng.random()` and, if drawn, `delta = rng.normal(0, std)`
    (genome.py:56-58), the cell of every spawn-fallback draw (ECO:759-762) — `env.rng` is wrapped in a
    recording proxy on the instance, the source is untouched;
  * the action dict of every step as (species, id, action) in the order it was passed (ECO:632);
  * the returned dicts, keyed (species, id) and sorted (the reference builds them from Python sets, so
    their order is hash-randomised, ECO:413,424): rewards, terminations, truncations, the float32
    observations (sha1 over all of them + the full arrays of a few steps), "__all__";
  * the state after the step: agent_positions (insertion order) / energies / ages / genome speeds,
    dead_prey, grass energies, sha1 of the float32 grid, active_num_* counters, self.agents.
The oracle (oracle/ppg_oracle_eco.c) must reproduce all of it bit for bit (tests/test_oracle_golden_eco.py).
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

ECO = "predpreygrass.evolutionary.eco_evolutionary"

CROWDED = dict(grid_size=7, initial_num_grass=20, n_initial_active_predators=6, n_initial_active_prey=14,
               predator_creation_energy_threshold=6.0, prey_creation_energy_threshold=3.5,
               initial_energy_predator=3.0, initial_energy_prey=2.0, energy_gain_per_step_grass=0.5,
               energy_loss_per_step_prey=0.02, energy_loss_per_step_predator=0.1,
               predator_obs_range=5, prey_obs_range=7, n_possible_predators=300, n_possible_prey=400)
RICH = dict(energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0, predator_creation_energy_threshold=8.0,
            energy_loss_per_step_predator=0.1)

CASES = [
    # name, overrides, seed, action order ("id" = predators then prey by ascending id, "shuffle"), max_calls
    ("eco_default_s1", {}, 1, "id", 400),
    ("eco_default_s2", {}, 2, "id", 400),
    ("eco_default_s5_shuffle", {}, 5, "shuffle", 400),
    ("eco_rich_s3", RICH, 3, "id", 160),
    ("eco_rich_s4_shuffle", dict(RICH, genome_mutation={"rate": 1.0, "std": 0.3}), 4, "shuffle", 160),
    ("eco_crowded_s1", CROWDED, 1, "id", 200),
    ("eco_crowded_s2_shuffle", dict(CROWDED, genome_mutation={"rate": 0.5, "std": 0.4}), 2, "shuffle", 200),
    ("eco_carcass_s1", dict(CROWDED, max_energy_gain_per_prey=0.8, max_energy_gain_per_grass=1.0), 1, "id", 200),
    ("eco_carcass_s6", dict(RICH, max_energy_gain_per_prey=1.5, max_energy_gain_per_grass=1.5, n_initial_active_predators=25,
                            n_initial_active_prey=40), 6, "shuffle", 120),
    ("eco_agecap_s2", dict(CROWDED, grid_size=10, max_agent_age={"predator": 25, "prey": 12}), 2, "id", 200),
    # age cap + finite intake cap: an aged-out prey bitten in the same step leaves a stale grid value behind (ECO:826-832)
    ("eco_ghost_s2", dict(CROWDED, grid_size=10, max_agent_age={"predator": 25, "prey": 12}, max_energy_gain_per_prey=1.0), 2, "id", 200),
    ("eco_juvenile_s3", dict(CROWDED, grid_size=10, carcass_only_predator_age={"predator": 6}, max_energy_gain_per_prey=0.7), 3, "id", 200),
    ("eco_trunc_s2", dict(RICH, max_steps=35), 2, "id", 60),
    ("eco_nogenome_s1", dict(RICH, genome_enabled=False, include_speed_in_obs=False), 1, "id", 120),
    ("eco_founders_fixed_s7", dict(RICH, founder_genome={"predator": {"speed_mean": 1.6, "speed_std": 0.0},
                                                         "prey": {"speed_mean": 1.2, "speed_std": 0.5}}), 7, "id", 250),
    # other window shapes (generic row writer on the device), odd grid, 3x3 and 7x7 action tables, other speed bounds /
    # threshold / jump distances (ECO:551-557,664-683)
    # lineage survival rewards (ECO:943-984): non-zero coefficients; births, deaths, carcasses, age-outs all move the counts
    ("eco_lineage_s1", dict(RICH, lineage_reward_coeff={"predator": 0.5, "prey": 0.25}), 1, "id", 160),
    ("eco_lineage_s2_shuffle", dict(CROWDED, grid_size=10, lineage_reward_coeff={"predator": 1.0, "prey": -0.125}, max_energy_gain_per_prey=0.8,
                                    genome_mutation={"rate": 0.5, "std": 0.4}), 2, "shuffle", 200),
    ("eco_lineage_age_s3", dict(CROWDED, grid_size=10, max_agent_age={"predator": 25, "prey": 12}, lineage_reward_coeff=0.75), 3, "id", 200),
    ("eco_narrow_s8", dict(RICH, grid_size=13, predator_obs_range=3, prey_obs_range=5, action_range=3, initial_num_grass=40), 8, "shuffle", 160),
    ("eco_jumps_s9", dict(RICH, grid_size=17, predator_obs_range=9, prey_obs_range=7, action_range=7, initial_num_grass=60,
                          trait_bounds={"speed": (0.25, 3.0)}, speed_distance_threshold=1.2, slow_max_move_distance=2,
                          fast_max_move_distance=3, founder_genome={"predator": {"speed_mean": 1.3, "speed_std": 0.6},
                                                                    "prey": {"speed_mean": 1.1, "speed_std": 0.5}}), 9, "id", 160),
]


def split(agent):
    kind, idx = str(agent).rsplit("_", 1)
    return (0 if kind == "predator" else 1), int(idx)


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


def record(name, overrides, seed, order, max_calls):
    mod = importlib.import_module(ECO + ".predpreygrass_rllib_env")
    cfgmod = importlib.import_module(ECO + ".config.config_env_eco_evolutionary")
    cfg = dict(cfgmod.config_env)
    cfg.update(overrides)
    env = mod.PredPreyGrass(cfg)
    G = env.grid_size

    obs, _ = env.reset(seed=seed)
    log = []
    env.rng = RecordingRng(env.rng, log)  # instrumentation of the instance, source untouched
    fallback = []
    orig_find = env._find_available_spawn_position

    def find_wrapped(ref_pos, occupied):
        n0 = len(log)
        pos = orig_find(ref_pos, occupied)
        drew = [e for e in log[n0:] if e[0] == "i"]
        del log[n0:]  # the index into a set-ordered list is not reproducible; the chosen cell is what counts
        if drew:
            fallback.append(int(pos[0]) * G + int(pos[1]))
        return pos

    env._find_available_spawn_position = find_wrapped

    founders = list(env.agents)
    founder_speed = [float(env.agent_genomes[a].speed) for a in founders] if env.genome_enabled else []
    init_cells = [int(env.agent_positions[a][0]) * G + int(env.agent_positions[a][1]) for a in founders]
    init_cells += [int(p[0]) * G + int(p[1]) for p in env.grass_positions.values()]
    reset_keys = sorted(split(a) for a in obs)
    reset_obs = [obs[f"{'predator' if s == 0 else 'prey'}_{i}"] for s, i in reset_keys]

    arng = np.random.default_rng(seed * 7919 + 13)
    n_act = env.action_range ** 2
    names = ("act_s", "act_id", "act_v", "row_s", "row_id", "row_rew", "row_term", "row_trunc", "st_s", "st_id", "st_x", "st_y",
             "st_e", "st_age", "st_speed", "st_dead", "ag_s", "ag_id")
    rec = {k: [] for k in names}
    offs = {k: [0] for k in ("act", "row", "st", "ag")}
    obs_sha, grid_sha, grass_e, all_term, all_trunc, steps, active = [], [], [], [], [], [], []
    full_obs = {}
    done = False
    calls = 0
    while not done and calls < max_calls:
        keys = sorted((a for a in env.agents), key=split)
        if order == "shuffle":
            keys = [keys[i] for i in arng.permutation(len(keys))]
        acts = {a: int(arng.integers(n_act)) for a in keys}
        obs, rew, term, trunc, _ = env.step(acts)
        ids = sorted((split(a) for a in obs))
        assert ids == sorted(split(a) for a in rew) == sorted(split(a) for a in term if a != "__all__") \
            == sorted(split(a) for a in trunc if a != "__all__"), (name, calls)
        for a, v in acts.items():
            s, i = split(a)
            rec["act_s"].append(s); rec["act_id"].append(i); rec["act_v"].append(v)
        offs["act"].append(len(rec["act_s"]))
        row_obs = []
        for s, i in ids:
            a = f"{'predator' if s == 0 else 'prey'}_{i}"
            rec["row_s"].append(s); rec["row_id"].append(i); rec["row_rew"].append(float(rew[a]))
            rec["row_term"].append(int(bool(term[a]))); rec["row_trunc"].append(int(bool(trunc[a])))
            assert obs[a].dtype == np.float32
            row_obs.append(obs[a])
        offs["row"].append(len(rec["row_s"]))
        obs_sha.append(sha(row_obs, np.float32))
        if calls < 2 or calls % 60 == 0:
            full_obs[calls] = row_obs
        pos_src = env.agent_positions
        for a, p in pos_src.items():
            s, i = split(a)
            rec["st_s"].append(s); rec["st_id"].append(i); rec["st_x"].append(int(p[0])); rec["st_y"].append(int(p[1]))
            rec["st_e"].append(float(env.agent_energies[a])); rec["st_age"].append(int(env.agent_ages[a]))
            g = env.agent_genomes.get(a)
            rec["st_speed"].append(float(g.speed) if g is not None else -1.0)
            rec["st_dead"].append(int(a in env.dead_prey))
        offs["st"].append(len(rec["st_s"]))
        for a in env.agents:
            s, i = split(a)
            rec["ag_s"].append(s); rec["ag_id"].append(i)
        offs["ag"].append(len(rec["ag_s"]))
        grid_sha.append(sha([env.grid_world_state], np.float32))
        assert env.grid_world_state.dtype == np.float32
        grass_e.append([float(env.grass_energies[g]) for g in env.grass_agents])
        all_term.append(int(bool(term["__all__"]))); all_trunc.append(int(bool(trunc["__all__"])))
        steps.append(env.current_step)
        active.append([int(env.active_num_predators), int(env.active_num_prey)])
        done = term["__all__"] or trunc["__all__"]
        calls += 1

    reals = [v for k, v in log if k in ("u", "n")]
    assert all(k in ("u", "n") for k, _ in log), "unexpected rng.integers outside the spawn fallback"
    out = dict(
        cfg_json=np.array(json.dumps(dict(cfg, variant="eco"), default=lambda o: None)), seed=np.int64(seed), order=np.array(order),
        init_cells=np.array(init_cells, np.int32), fallback_cells=np.array(fallback, np.int32),
        founder_speed=np.array(founder_speed, np.float64), step_reals=np.array(reals, np.float64),
        reset_row_s=np.array([k[0] for k in reset_keys], np.int8), reset_row_id=np.array([k[1] for k in reset_keys], np.int32),
        reset_sha=sha(reset_obs, np.float32),
        act_off=np.array(offs["act"], np.int64), row_off=np.array(offs["row"], np.int64),
        st_off=np.array(offs["st"], np.int64), ag_off=np.array(offs["ag"], np.int64),
        act_s=np.array(rec["act_s"], np.int8), act_id=np.array(rec["act_id"], np.int32), act_v=np.array(rec["act_v"], np.int8),
        row_s=np.array(rec["row_s"], np.int8), row_id=np.array(rec["row_id"], np.int32),
        row_rew=np.array(rec["row_rew"], np.float64), row_term=np.array(rec["row_term"], np.int8),
        row_trunc=np.array(rec["row_trunc"], np.int8),
        st_s=np.array(rec["st_s"], np.int8), st_id=np.array(rec["st_id"], np.int32), st_x=np.array(rec["st_x"], np.int16),
        st_y=np.array(rec["st_y"], np.int16), st_e=np.array(rec["st_e"], np.float64), st_age=np.array(rec["st_age"], np.int32),
        st_speed=np.array(rec["st_speed"], np.float64), st_dead=np.array(rec["st_dead"], np.int8),
        ag_s=np.array(rec["ag_s"], np.int8), ag_id=np.array(rec["ag_id"], np.int32),
        obs_sha=np.array(obs_sha, np.uint8).reshape(-1, 20), grid_sha=np.array(grid_sha, np.uint8).reshape(-1, 20),
        grass_e=np.array(grass_e, np.float64), all_term=np.array(all_term, np.int8), all_trunc=np.array(all_trunc, np.int8),
        steps=np.array(steps, np.int32), active=np.array(active, np.int32).reshape(-1, 2),
        full_obs_steps=np.array(sorted(full_obs), np.int32),
    )
    for t, rows in full_obs.items():
        for s in range(2):
            sel = [r for r in rows if r.shape[-1] == (env.predator_obs_range if s == 0 else env.prey_obs_range)]
            if env.predator_obs_range == env.prey_obs_range:
                raise SystemExit("cases need distinct obs ranges to split the full-obs dump")
            out[f"full_obs_{t}_{s}"] = np.stack(sel) if sel else np.zeros((0,), np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    births = int(env.spawned_predators + env.spawned_prey)
    n_mut = sum(1 for k, _ in log if k == "n")
    print(f"{name:26s} calls={calls:4d} births={births:4d} mutations={n_mut:3d} fallback={len(fallback):3d} "
          f"end={'term' if all_term[-1] else ('trunc' if all_trunc[-1] else 'cut')} size={os.path.getsize(path)//1024}KB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        record(*case)
