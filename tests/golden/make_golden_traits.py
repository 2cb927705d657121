#!/usr/bin/env python
"""Record golden trajectories from the UNMODIFIED reference classes of the other heritable-trait variants:
  mr_*    predpreygrass/evolutionary/eco_evolutionary_metabolic_rate/predpreygrass_rllib_env.py   (MR)
  inv_*   predpreygrass/evolutionary/eco_evolutionary_investment/predpreygrass_rllib_env.py       (INV)
  coop_*  predpreygrass/evolutionary/eco_evolutionary_cooperation/predpreygrass_rllib_env.py      (COOP)
  cad_*   predpreygrass/evolutionary/eco_evolutionary_cadence/predpreygrass_rllib_env.py          (CAD; its observations
          are dicts {"observations", "action_mask"}: the window is recorded like the others', the mask as one "frozen" bit per
          row after checking that it is all ones or "stay" only; also the founders' move accumulators, the accumulators
          after every step and, per birth, the extra `rng.uniform` phase draw, CAD:1324-1327)

Runs in the build container only (needs /root/reference); the GPU box never executes this.
Usage:  python tests/golden/make_golden_traits.py [case ...]

Same method as make_golden_eco.py: the class is imported under the stub `ray` / `gymnasium` packages of
tests/golden/_shim and stepped with seeded uniform-random actions.  Recorded per case:
  * the RNG tape: the number of founders the reset drew (MR:189-192), the founders' trait values (`agent_genomes` after
    reset, genome.py `founder_genome`), the initial cells (`rng.choice`, MR:1469), per birth `u = rng.random()` and, if
    drawn, `delta = rng.normal(0, std)` (genome.py `mutate_genome`), the cell of every spawn-fallback draw (MR:700-712);
  * the action dict of every step in the order it was passed (MR:581);
  * the returned dicts keyed (species, id) and sorted (the reference builds them from Python sets): rewards,
    terminations, truncations, the float32 observations (sha1 + full arrays of a few steps), "__all__";
  * the state after the step: agent_positions / energies / ages / trait values, grass energies, sha1 of the float32
    grid, active_num_* counters, self.agents; and `infos["__all__"]["training_metrics"]` of the episode's last step.
The oracle (oracle/ppg_oracle_eco.c, trait_mode branches) must reproduce all of it bit for bit
(tests/test_oracle_golden_traits.py).
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

PKG = {"mr": "predpreygrass.evolutionary.eco_evolutionary_metabolic_rate",
       "inv": "predpreygrass.evolutionary.eco_evolutionary_investment",
       "coop": "predpreygrass.evolutionary.eco_evolutionary_cooperation",
       "cad": "predpreygrass.evolutionary.eco_evolutionary_cadence"}
TRAIT = {"mr": "metabolic_rate", "inv": "offspring_investment_fraction", "coop": "cooperation_rate", "cad": "speed"}

CROWDED = dict(grid_size=9, initial_num_grass=24, n_initial_active_predators=6, n_initial_active_prey=14,
               predator_creation_energy_threshold=6.0, prey_creation_energy_threshold=4.5, energy_gain_per_step_grass=0.3,
               predator_obs_range=5, prey_obs_range=7, n_possible_predators=300, n_possible_prey=400)
RICH = dict(energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0, predator_creation_energy_threshold=8.0)


CASES = [
    # name, overrides, seed, action order, max_calls
    ("mr_default_s1", {}, 1, "id", 400),
    ("mr_default_s2_shuffle", {}, 2, "shuffle", 400),
    ("mr_rich_s3", dict(RICH, basal_energy_cost_predator=0.1), 3, "id", 160),
    ("mr_rich_mut_s4_shuffle", dict(RICH, basal_energy_cost_predator=0.1, genome_mutation={"rate": 1.0, "std": 0.3},
                                    metabolic_rate_alpha=0.7, predator_satiation_cooldown=3), 4, "shuffle", 160),
    ("mr_crowded_s1", dict(CROWDED, basal_energy_cost_prey=0.02, basal_energy_cost_predator=0.1, initial_energy_predator=3.0,
                           initial_energy_prey=2.0, max_energy_gain_per_prey=1.5, movement_energy_cost_per_cell_prey=0.03,
                           movement_energy_cost_per_cell_predator=0.05), 1, "id", 200),
    ("mr_density_s5", dict(CROWDED, basal_energy_cost_prey=0.02, basal_energy_cost_predator=0.05, initial_energy_predator=4.0,
                           predator_reproduction_max_ratio=0.4, predator_satiation_cooldown=0, genome_mutation={"rate": 0.5, "std": 0.2},
                           n_initial_active_predators_min=6, n_initial_active_prey_min=10), 5, "shuffle", 200),
    ("mr_trunc_s2", dict(RICH, max_steps=35), 2, "id", 60),
    ("mr_nogenome_s1", dict(RICH, genome_enabled=False), 1, "id", 120),
    ("inv_default_s1", {}, 1, "id", 400),
    ("inv_default_s3_shuffle", {}, 3, "shuffle", 400),
    ("inv_rich_s2", dict(RICH, energy_loss_per_step_predator=0.1, genome_mutation={"rate": 0.6, "std": 0.15}), 2, "id", 130),
    ("inv_crowded_s4", dict(CROWDED, energy_loss_per_step_prey=0.02, energy_loss_per_step_predator=0.1,
                            initial_energy_predator_at_reset=3.0, initial_energy_prey_at_reset=2.0, max_energy_gain_per_prey=2.0,
                            predator_satiation_cooldown=2), 4, "shuffle", 200),
    ("inv_nogenome_s6", dict(RICH, genome_enabled=False, energy_loss_per_step_predator=0.1), 6, "id", 120),
    ("coop_default_s1", {}, 1, "id", 400),
    ("coop_share_s2", dict(RICH, basal_energy_cost_predator=0.1, cooperation_range=3,
                           founder_genome={"predator": {"cooperation_rate_mean": 0.4, "cooperation_rate_std": 0.2},
                                           "prey": {"cooperation_rate_mean": 0.3, "cooperation_rate_std": 0.2}}), 2, "shuffle", 160),
    ("coop_crowded_s3", dict(CROWDED, basal_energy_cost_prey=0.02, basal_energy_cost_predator=0.1, initial_energy_predator=3.0,
                             initial_energy_prey=2.0, cooperation_range=1, genome_mutation={"rate": 0.5, "std": 0.3},
                             founder_genome={"predator": {"cooperation_rate_mean": 0.5, "cooperation_rate_std": 0.3},
                                             "prey": {"cooperation_rate_mean": 0.5, "cooperation_rate_std": 0.3}}), 3, "id", 200),
    ("coop_trunc_s4", dict(RICH, max_steps=30, founder_genome={"predator": {"cooperation_rate_mean": 0.2, "cooperation_rate_std": 0.1},
                                                               "prey": {"cooperation_rate_mean": 0.2, "cooperation_rate_std": 0.1}}), 4, "id", 60),
    ("cad_default_s1", {}, 1, "id", 300),
    ("cad_default_s2_shuffle", {}, 2, "shuffle", 300),
    ("cad_rich_s3", dict(RICH, energy_loss_per_step_predator=0.1, genome_mutation={"rate": 0.6, "std": 0.15}), 3, "id", 140),
    ("cad_crowded_s4", dict(CROWDED, energy_loss_per_step_prey=0.02, energy_loss_per_step_predator=0.1, initial_energy_predator=3.0,
                            initial_energy_prey=2.0, max_cooldown=3, genome_mutation={"rate": 0.5, "std": 0.3}), 4, "shuffle", 200),
    ("cad_aging_s5", dict(RICH, max_agent_age={"predator": 30, "prey": 12}, max_cooldown=4, metabolic_speed_coeff=0.0,
                          max_energy_gain_per_grass=1.0), 5, "id", 120),
    ("cad_nogenome_s6", dict(RICH, genome_enabled=False, include_speed_in_obs=False, energy_loss_per_step_predator=0.1), 6, "id", 120),
    ("cad_trunc_s7", dict(RICH, max_steps=30), 7, "shuffle", 60),
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

    def uniform(self, *a, **k):
        v = self._rng.uniform(*a, **k)
        self._log.append(("f", float(v)))
        return v

    def integers(self, *a, **k):
        v = self._rng.integers(*a, **k)
        self._log.append(("i", int(v)))
        return v

    def __getattr__(self, name):
        return getattr(self._rng, name)


def record(name, overrides, seed, order, max_calls):
    fam = name.split("_")[0]
    trait = TRAIT[fam]
    mod = importlib.import_module(PKG[fam] + ".predpreygrass_rllib_env")
    cfgmod = importlib.import_module(PKG[fam] + ".config.config_env_eco_evolutionary")
    cfg = dict(cfgmod.config_env)
    cfg.update(overrides)
    env = mod.PredPreyGrass(cfg)
    G = env.grid_size
    NAMES = ("predator", "prey")

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

    cad = fam == "cad"
    win = (lambda o: o["observations"]) if cad else (lambda o: o)

    def frozen(o):  # CAD:746-753: the mask is all ones, or "stay" only
        if not cad:
            return 0
        m = np.asarray(o["action_mask"])
        assert m.dtype == np.float32 and m.shape == (n_act_total,)
        stay_only = np.zeros(n_act_total, np.float32); stay_only[n_act_total // 2] = 1.0
        assert np.array_equal(m, np.ones(n_act_total, np.float32)) or np.array_equal(m, stay_only), m
        return int(m.sum() == 1.0)

    n_act_total = env.action_range ** 2
    founders = list(env.agents)
    n_found = [sum(1 for a in founders if split(a)[0] == s) for s in range(2)]
    founder_trait = [float(getattr(env.agent_genomes[a], trait)) for a in founders] if env.genome_enabled else []
    founder_acc = [float(env.agent_move_accumulator[a]) for a in founders] if cad else []
    init_cells = [int(env.agent_positions[a][0]) * G + int(env.agent_positions[a][1]) for a in founders]
    init_cells += [int(p[0]) * G + int(p[1]) for p in env.grass_positions.values()]
    reset_keys = sorted(split(a) for a in obs)
    reset_obs = [win(obs[f"{NAMES[s]}_{i}"]) for s, i in reset_keys]
    reset_frozen = [frozen(obs[f"{NAMES[s]}_{i}"]) for s, i in reset_keys]

    arng = np.random.default_rng(seed * 7919 + 13)
    n_act = env.action_range ** 2
    names = ("act_s", "act_id", "act_v", "row_s", "row_id", "row_rew", "row_term", "row_trunc", "row_frozen", "st_s", "st_id", "st_x", "st_y",
             "st_e", "st_age", "st_trait", "st_acc", "ag_s", "ag_id")
    rec = {k: [] for k in names}
    offs = {k: [0] for k in ("act", "row", "st", "ag")}
    obs_sha, grid_sha, grass_e, all_term, all_trunc, steps, active = [], [], [], [], [], [], []
    full_obs = {}
    metrics = None
    done = False
    calls = 0
    while not done and calls < max_calls:
        keys = sorted((a for a in env.agents), key=split)
        if order == "shuffle":
            keys = [keys[i] for i in arng.permutation(len(keys))]
        acts = {a: int(arng.integers(n_act)) for a in keys}
        try:
            obs, rew, term, trunc, infos = env.step(acts)
        except RuntimeError as ex:  # "No free spawn position available" (MR:895,979): the world is full, the recording ends before this call
            print(f"  {name}: reference raised {ex!r} at call {calls}; recording cut there")
            break
        ids = sorted((split(a) for a in obs))
        assert ids == sorted(split(a) for a in rew) == sorted(split(a) for a in term if a != "__all__") \
            == sorted(split(a) for a in trunc if a != "__all__"), (name, calls)
        for a, v in acts.items():
            s, i = split(a)
            rec["act_s"].append(s); rec["act_id"].append(i); rec["act_v"].append(v)
        offs["act"].append(len(rec["act_s"]))
        row_obs = []
        for s, i in ids:
            a = f"{NAMES[s]}_{i}"
            rec["row_s"].append(s); rec["row_id"].append(i); rec["row_rew"].append(float(rew[a]))
            rec["row_term"].append(int(bool(term[a]))); rec["row_trunc"].append(int(bool(trunc[a])))
            assert win(obs[a]).dtype == np.float32
            row_obs.append(win(obs[a]))
            rec["row_frozen"].append(frozen(obs[a]))
        offs["row"].append(len(rec["row_s"]))
        obs_sha.append(sha(row_obs, np.float32))
        if calls < 2 or calls % 60 == 0:
            full_obs[calls] = row_obs
        for a, p in env.agent_positions.items():
            s, i = split(a)
            rec["st_s"].append(s); rec["st_id"].append(i); rec["st_x"].append(int(p[0])); rec["st_y"].append(int(p[1]))
            rec["st_e"].append(float(env.agent_energies[a])); rec["st_age"].append(int(env.agent_ages[a]))
            g = env.agent_genomes.get(a)
            rec["st_trait"].append(float(getattr(g, trait)) if g is not None else -1.0)
            rec["st_acc"].append(float(env.agent_move_accumulator[a]) if cad else 0.0)
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
        if done:
            metrics = {k: float(v) for k, v in infos["__all__"]["training_metrics"].items()}
        calls += 1

    reals = [v for k, v in log if k in ("u", "n", "f")]
    assert all(k in ("u", "n", "f") for k, _ in log), "unexpected rng.integers outside the spawn search"
    out = dict(
        cfg_json=np.array(json.dumps(dict(cfg, variant=fam), default=lambda o: None)), seed=np.int64(seed), order=np.array(order),
        n_found=np.array(n_found, np.int32),
        init_cells=np.array(init_cells, np.int32), fallback_cells=np.array(fallback, np.int32),
        founder_trait=np.array(founder_trait, np.float64), founder_acc=np.array(founder_acc, np.float64),
        step_reals=np.array(reals, np.float64), reset_frozen=np.array(reset_frozen, np.int8),
        row_frozen=np.array(rec["row_frozen"], np.int8), st_acc=np.array(rec["st_acc"], np.float64),
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
        st_trait=np.array(rec["st_trait"], np.float64),
        ag_s=np.array(rec["ag_s"], np.int8), ag_id=np.array(rec["ag_id"], np.int32),
        obs_sha=np.array(obs_sha, np.uint8).reshape(-1, 20), grid_sha=np.array(grid_sha, np.uint8).reshape(-1, 20),
        grass_e=np.array(grass_e, np.float64), all_term=np.array(all_term, np.int8), all_trunc=np.array(all_trunc, np.int8),
        steps=np.array(steps, np.int32), active=np.array(active, np.int32).reshape(-1, 2),
        full_obs_steps=np.array(sorted(full_obs), np.int32),
        metrics_json=np.array(json.dumps(metrics, sort_keys=True)),
    )
    for t, rows in full_obs.items():
        for s in range(2):
            sel = [r for r in rows if r.shape[-1] == (env.predator_obs_range if s == 0 else env.prey_obs_range)]
            out[f"full_obs_{t}_{s}"] = np.stack(sel) if sel else np.zeros((0,), np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    births = int(env.spawned_predators + env.spawned_prey)
    n_mut = sum(1 for k, _ in log if k == "n")
    print(f"{name:26s} founders={n_found} calls={calls:4d} births={births:4d} mutations={n_mut:3d} fallback={len(fallback):3d} "
          f"end={'term' if all_term[-1] else ('trunc' if all_trunc[-1] else 'cut')} size={os.path.getsize(path)//1024}KB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        record(*case)
