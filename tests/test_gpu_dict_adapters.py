"""`PredPreyGrassEco` / `PredPreyGrassStag` (the reference's MultiAgentEnv dict API over the CUDA step) against the
episodes recorded from the unmodified reference classes through that same API (`-m gpu`).  reset(seed) uses the
adapter's own host-side reproduction of the reference's reset draws; the draws after the reset (mutation, spawn
fallback, capture success) are the reference's recorded ones (`options={"ppg_tape": ...}`)."""
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN_DIR, golden_cases, load_golden, sha_f32

pytestmark = pytest.mark.gpu


def _caps(cfg, keys):
    return tuple(min(sum(cfg.get(k, 0) for k in ks), lim) for ks, lim in zip(keys, (224, 416)))


@pytest.mark.parametrize("name", golden_cases(("eco",)))
def test_eco_dict_adapter_replays_reference_episode(name):
    from predpreygrass_b200.env_evolutionary import PredPreyGrassEco

    z, cfg = load_golden(name)
    cfg.pop("variant")
    cfg["cap_live"] = _caps(cfg, (("n_possible_predators",), ("n_possible_prey",)))
    env = PredPreyGrassEco(cfg)
    names = ("predator", "prey")
    key = lambda s, i: f"{names[s]}_{i}"  # noqa: E731
    obs, infos = env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    assert infos == {}
    assert list(obs) == [key(s, i) for s, i in zip(z["reset_row_s"], z["reset_row_id"])]
    assert np.array_equal(sha_f32([obs[k] for k in obs]), z["reset_sha"])
    for a, o in obs.items():
        assert o.dtype == np.float32 and o.shape == env.observation_spaces[a].shape
    shuffle = str(z["order"]) == "shuffle"
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        acts = {key(s, i): int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
        if not shuffle:  # the recording passed the actions in id order; the adapter's row order is the same
            assert list(acts) == sorted(acts, key=lambda a: (a[:4] == "prey", int(a.rsplit("_", 1)[1])))
        obs, rew, term, trunc, infos = env.step(acts)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        keys = [key(s, i) for s, i in zip(z["row_s"][r0:r1], z["row_id"][r0:r1])]
        assert list(obs) == keys and list(rew) == keys, (name, t)
        assert np.array_equal(np.array([rew[k] for k in keys], np.float32), z["row_rew"][r0:r1].astype(np.float32)), (name, t)
        assert [int(term[k]) for k in keys] == list(z["row_term"][r0:r1]), (name, t)
        assert [int(trunc[k]) for k in keys] == list(z["row_trunc"][r0:r1]), (name, t)
        assert np.array_equal(sha_f32([obs[k] for k in keys]), z["obs_sha"][t]), (name, t)
        assert term["__all__"] == bool(z["all_term"][t]) and trunc["__all__"] == bool(z["all_trunc"][t]), (name, t)
        assert set(infos) - {"__all__"} == set(keys)
        assert env.current_step == int(z["steps"][t])
        if term["__all__"] or trunc["__all__"]:
            assert env.agents == []
            # infos["__all__"]["training_metrics"] = `_build_episode_training_metrics` (ECO:1613-1668) of the reference's own run
            # of this episode (tests/golden/make_golden_eco_metrics.py): same keys in the same order; counts exact, the
            # float means to 1e-9 (the device keeps per-species totals, the reference averages per-agent sums)
            want = json.load(open(os.path.join(GOLDEN_DIR, "eco_training_metrics.json")))[name]
            got = infos["__all__"]["training_metrics"]
            assert sorted(got) == sorted(want), (name, sorted(set(got) ^ set(want)))
            for k, v in want.items():
                if k.endswith("_count") or k.endswith("_fraction_fast") or "_speed_" in k:
                    assert got[k] == pytest.approx(v, rel=1e-12, abs=0.0), (name, k, got[k], v)
                else:
                    assert got[k] == pytest.approx(v, rel=1e-9, abs=1e-12), (name, k, got[k], v)
            break
        g0, g1 = z["ag_off"][t], z["ag_off"][t + 1]
        assert env.agents == [key(s, i) for s, i in zip(z["ag_s"][g0:g1], z["ag_id"][g0:g1])], (name, t)
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        want = {key(s, i): ((int(x), int(y)), float(e), int(a), float(v)) for s, i, x, y, e, a, v in
                zip(z["st_s"][s0:s1], z["st_id"][s0:s1], z["st_x"][s0:s1], z["st_y"][s0:s1], z["st_e"][s0:s1],
                    z["st_age"][s0:s1], z["st_speed"][s0:s1])}
        pos, en, age, sp = env.agent_positions, env.agent_energies, env.agent_ages, env.agent_speeds
        assert sorted(pos) == sorted(want), (name, t)
        assert all((pos[k], en[k], age[k], sp[k]) == want[k] for k in want), (name, t)
        assert (env.active_num_predators, env.active_num_prey) == tuple(z["active"][t])
    with pytest.raises(KeyError):
        env.reset(seed=1)
        env.step({"predator_399": 0})
    env.close()


@pytest.mark.parametrize("name", golden_cases(("stag",)))
def test_stag_dict_adapter_replays_reference_episode(name):
    from predpreygrass_b200.env_evolutionary import FACING_OPTIONS, PredPreyGrassStag

    z, cfg = load_golden(name)
    cfg.pop("variant")
    cfg["cap_live"] = _caps(cfg, (("n_possible_type_1_predators", "n_possible_type_2_predators"),
                                  ("n_possible_type_1_prey", "n_possible_type_2_prey")))
    env = PredPreyGrassStag(cfg)
    n1 = (cfg.get("n_possible_type_1_predators", 0), cfg.get("n_possible_type_1_prey", 0))
    sp = ("predator", "prey")
    key = lambda s, i: f"type_1_{sp[s]}_{i}" if i < n1[s] else f"type_2_{sp[s]}_{i - n1[s]}"  # noqa: E731
    obs, infos = env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["step_ints"], z["step_reals"])})
    assert list(obs) == [key(s, i) for s, i in zip(z["reset_row_s"], z["reset_row_id"])] == env.agents
    assert np.array_equal(sha_f32([obs[k] for k in obs]), z["reset_sha"])
    rng = np.random.default_rng(0)
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        acts = {}
        for s, i, mv, jn in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_move"][a0:a1], z["act_join"][a0:a1]):
            if s == 1:
                acts[key(s, i)] = int(mv)
            else:  # the three accepted predator action forms (STAG:771-799)
                form = rng.integers(3)
                acts[key(s, i)] = (np.array([mv, jn]) if form == 0 else (int(mv), int(jn)) if form == 1
                                   else {"move": int(mv), "join_hunt": int(jn)})
        assert list(acts) == env.agents or str(z["order"]) == "shuffle", (name, t)  # the recording acted for self.agents
        obs, rew, term, trunc, infos = env.step(acts)
        g0, g1 = z["ag_off"][t], z["ag_off"][t + 1]
        agents = [key(s, i) for s, i in zip(z["ag_s"][g0:g1], z["ag_id"][g0:g1])]
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        keys = [key(s, i) for s, i in zip(z["row_s"][r0:r1], z["row_id"][r0:r1])]
        live = [k for k in agents if k in obs and not term[k]]
        assert [k for k in obs if not term[k]] == live, (name, t)   # observation-dict order of the live agents (STAG:551)
        assert set(obs) == set(keys) and set(rew) == set(keys), (name, t)
        assert np.array_equal(np.array([rew[k] for k in keys], np.float32), z["row_rew"][r0:r1].astype(np.float32)), (name, t)
        assert [int(term[k]) for k in keys] == list(z["row_term"][r0:r1]), (name, t)
        assert [int(trunc[k]) for k in keys] == list(z["row_trunc"][r0:r1]), (name, t)
        ordered = [obs[k] for k in keys if "predator" in k] + [obs[k] for k in keys if "prey" in k]
        assert np.array_equal(sha_f32(ordered), z["obs_sha"][t]), (name, t)
        assert term["__all__"] == bool(z["all_term"][t]) and trunc["__all__"] == bool(z["all_trunc"][t]), (name, t)
        assert env.current_step == int(z["steps"][t])
        # the recording kept the positioned part of `env.agents` (under strict_rllib_output the list also carries the
        # ids that ended this step, STAG:565-573)
        assert [a for a in env.agents if not term.get(a, False)] == agents, (name, t)
        if cfg.get("strict_rllib_output", False):
            assert sorted(env.agents) == sorted(keys), (name, t)
        c = z["counters"][t]
        assert infos["__all__"]["team_capture_successes"] == c[0] and infos["__all__"]["team_capture_attempts"] == c[8], (name, t)
        if term["__all__"] or trunc["__all__"]:
            break
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        want = {key(s, i): ((int(x), int(y)), float(e), int(a)) for s, i, x, y, e, a in
                zip(z["st_s"][s0:s1], z["st_id"][s0:s1], z["st_x"][s0:s1], z["st_y"][s0:s1], z["st_e"][s0:s1], z["st_age"][s0:s1])}
        pos, en, age = env.agent_positions, env.agent_energies, env.agent_ages
        assert list(pos) != [] and sorted(pos) == sorted(want), (name, t)
        assert all((pos[k], en[k], age[k]) == want[k] for k in want), (name, t)
        # get_total_energy_by_type (STAG:1964-1997), the reference's sums over its own dict order
        tot = {"predator": 0.0, "prey": 0.0, "grass": sum(float(e) for e in z["grass_e"][t]), "type_1_predator": 0.0, "type_2_predator": 0.0,
               "type_1_prey": 0.0, "type_2_prey": 0.0}
        for s, i, e in zip(z["st_s"][s0:s1], z["st_id"][s0:s1], z["st_e"][s0:s1]):
            k = key(s, i)
            role = "predator" if "predator" in k else "prey"
            tot[role] += float(e)
            tot[k[:7] + role] += float(e)
        assert env.get_total_energy_by_type() == pytest.approx(tot, rel=1e-12, abs=0.0), (name, t)
        face, trait = env.predator_facing, env.predator_cooperation_trait
        for s, i, f, v in zip(z["st_s"][s0:s1], z["st_id"][s0:s1], z["st_face"][s0:s1], z["st_trait"][s0:s1]):
            if s == 0:
                assert face[key(s, i)] == FACING_OPTIONS[int(f)] and trait[key(s, i)] == float(v), (name, t)
    env.close()


def test_eco_adapter_time_limit_reports_training_metrics():
    """reference test_time_limit_truncates_with_final_bootstrap_observations (eco tests :296-329) through the dict API:
    max_steps = 1 -> everybody truncated, `agents` cleared, `infos["__all__"]["training_metrics"]` carries the keys of
    `_build_episode_training_metrics` (ECO:1613-1661), the ones `utils/episode_return_callback.py:9-80` reads"""
    from predpreygrass_b200.config import ECO_CONFIG
    from predpreygrass_b200.env_evolutionary import PredPreyGrassEco

    cfg = dict(ECO_CONFIG, max_steps=1, n_initial_active_predators=1, n_initial_active_prey=1, n_possible_predators=8, n_possible_prey=8,
               initial_num_grass=4, predator_creation_energy_threshold=999.0, prey_creation_energy_threshold=999.0, cap_live=(32, 32))
    env = PredPreyGrassEco(cfg)
    obs, _ = env.reset(seed=125)
    assert set(obs) == {"predator_0", "prey_0"} and env.action_spaces["predator_0"].n == 25
    stay = next(a for a, mv in env.action_to_move_tuple_agents.items() if mv == (0, 0))
    obs, rew, term, trunc, infos = env.step({a: stay for a in env.agents})
    assert set(obs) == {"predator_0", "prey_0"}
    assert obs["predator_0"].shape == (4, 7, 7) and obs["prey_0"].shape == (4, 9, 9)
    assert all(trunc[a] and not term[a] for a in obs) and trunc["__all__"] and not term["__all__"]
    assert env.agents == []
    m = infos["__all__"]["training_metrics"]
    for role in ("predator", "prey"):
        for k in ("speed_mean", "speed_std", "speed_p25", "speed_p50", "speed_p75", "fraction_fast", "distance_traveled_mean",
                  "movement_energy_spent_mean", "offspring_count_mean", "agent_count"):
            assert f"{role}_{k}" in m
    assert len(m) == 20 and m["predator_agent_count"] == 1.0 and 0.5 <= m["predator_speed_mean"] <= 2.0
    assert m["prey_distance_traveled_mean"] == 0.0 and m["prey_offspring_count_mean"] == 0.0
    env.close()


@pytest.mark.parametrize("name", golden_cases(("mr", "inv", "coop")))
def test_trait_dict_adapter_replays_reference_episode(name):
    """PredPreyGrassMetabolicRate / Investment / Cooperation through the reference's dict API: reset(seed) with the adapter's
    own reproduction of the reset draws (founder counts included), then the recorded episode; at the episode's end the
    `training_metrics` of the reference's own run (stored in the golden file) for every key the adapter emits."""
    from predpreygrass_b200 import env_evolutionary as E

    z, cfg = load_golden(name)
    cls = {"mr": E.PredPreyGrassMetabolicRate, "inv": E.PredPreyGrassInvestment, "coop": E.PredPreyGrassCooperation}[cfg.pop("variant")]
    cfg["cap_live"] = _caps(cfg, (("n_possible_predators",), ("n_possible_prey",)))
    env = cls(cfg)
    names = ("predator", "prey")
    key = lambda s, i: f"{names[s]}_{i}"  # noqa: E731
    obs, infos = env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    assert infos == {}
    assert list(obs) == [key(s, i) for s, i in zip(z["reset_row_s"], z["reset_row_id"])]
    assert np.array_equal(sha_f32([obs[k] for k in obs]), z["reset_sha"])
    for a, o in obs.items():
        assert o.dtype == np.float32 and o.shape == env.observation_spaces[a].shape
    ended = False
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        acts = {key(s, i): int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
        obs, rew, term, trunc, infos = env.step(acts)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        keys = [key(s, i) for s, i in zip(z["row_s"][r0:r1], z["row_id"][r0:r1])]
        assert list(obs) == keys and list(rew) == keys, (name, t)
        assert np.array_equal(np.array([rew[k] for k in keys], np.float32), z["row_rew"][r0:r1].astype(np.float32)), (name, t)
        assert [int(term[k]) for k in keys] == list(z["row_term"][r0:r1]), (name, t)
        assert [int(trunc[k]) for k in keys] == list(z["row_trunc"][r0:r1]), (name, t)
        assert np.array_equal(sha_f32([obs[k] for k in keys]), z["obs_sha"][t]), (name, t)
        assert term["__all__"] == bool(z["all_term"][t]) and trunc["__all__"] == bool(z["all_trunc"][t]), (name, t)
        assert env.current_step == int(z["steps"][t])
        if term["__all__"] or trunc["__all__"]:
            ended = True
            assert env.agents == []
            want = json.loads(str(z["metrics_json"]))
            got = infos["__all__"]["training_metrics"]
            assert set(got) <= set(want), (name, sorted(set(got) - set(want)))
            missing = {k for k in want if k not in got}
            assert not missing, (name, sorted(missing))  # every key of the reference's `_build_episode_training_metrics`
            if "predator_energy_donated_total" in got:  # the device's donation totals against the host replay's
                assert got["predator_energy_donated_total"] == env._events.energy_donated[0] and got["prey_energy_donated_total"] == env._events.energy_donated[1]
            for k, v in got.items():
                assert v == pytest.approx(want[k], rel=1e-9, abs=1e-12), (name, k, v, want[k])
            break
        g0, g1 = z["ag_off"][t], z["ag_off"][t + 1]
        assert env.agents == [key(s, i) for s, i in zip(z["ag_s"][g0:g1], z["ag_id"][g0:g1])], (name, t)
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        want = {key(s, i): ((int(x), int(y)), float(e), int(a), float(v)) for s, i, x, y, e, a, v in
                zip(z["st_s"][s0:s1], z["st_id"][s0:s1], z["st_x"][s0:s1], z["st_y"][s0:s1], z["st_e"][s0:s1],
                    z["st_age"][s0:s1], z["st_trait"][s0:s1])}
        pos, en, age, tr = env.agent_positions, env.agent_energies, env.agent_ages, env.agent_traits
        assert sorted(pos) == sorted(want), (name, t)
        assert all((pos[k], en[k], age[k], tr[k]) == want[k] for k in want), (name, t)
        assert (env.active_num_predators, env.active_num_prey) == tuple(z["active"][t])
    assert ended or str(z["metrics_json"]) == "null"
    env.close()


@pytest.mark.parametrize("name", golden_cases(("cad",)))
def test_cadence_dict_adapter_replays_reference_episode(name):
    """PredPreyGrassCadence through the reference's dict API: observation dicts {"observations", "action_mask"}, the
    accumulators after every step, the episode's training_metrics."""
    from predpreygrass_b200.env_evolutionary import PredPreyGrassCadence

    z, cfg = load_golden(name)
    cfg.pop("variant")
    cfg["cap_live"] = _caps(cfg, (("n_possible_predators",), ("n_possible_prey",)))
    env = PredPreyGrassCadence(cfg)
    names = ("predator", "prey")
    key = lambda s, i: f"{names[s]}_{i}"  # noqa: E731
    n_act = cfg["action_range"] ** 2

    def check_masks(obs, keys, frozen):
        for k, fr in zip(keys, frozen):
            m = obs[k]["action_mask"]
            assert m.dtype == np.float32 and m.shape == (n_act,)
            assert (m.sum() == 1.0 and m[n_act // 2] == 1.0) if fr else bool((m == 1.0).all()), (name, k, m)

    obs, infos = env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    keys = [key(s, i) for s, i in zip(z["reset_row_s"], z["reset_row_id"])]
    assert list(obs) == keys
    assert np.array_equal(sha_f32([obs[k]["observations"] for k in obs]), z["reset_sha"])
    check_masks(obs, keys, z["reset_frozen"])
    for a, o in obs.items():
        assert o["observations"].shape == env.observation_spaces[a]["observations"].shape
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        acts = {key(s, i): int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
        obs, rew, term, trunc, infos = env.step(acts)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        keys = [key(s, i) for s, i in zip(z["row_s"][r0:r1], z["row_id"][r0:r1])]
        assert list(obs) == keys and list(rew) == keys, (name, t)
        assert np.array_equal(np.array([rew[k] for k in keys], np.float32), z["row_rew"][r0:r1].astype(np.float32)), (name, t)
        assert [int(term[k]) for k in keys] == list(z["row_term"][r0:r1]), (name, t)
        assert [int(trunc[k]) for k in keys] == list(z["row_trunc"][r0:r1]), (name, t)
        assert np.array_equal(sha_f32([obs[k]["observations"] for k in keys]), z["obs_sha"][t]), (name, t)
        check_masks(obs, keys, z["row_frozen"][r0:r1])
        assert term["__all__"] == bool(z["all_term"][t]) and trunc["__all__"] == bool(z["all_trunc"][t]), (name, t)
        if term["__all__"] or trunc["__all__"]:
            want = json.loads(str(z["metrics_json"]))
            got = infos["__all__"]["training_metrics"]
            assert sorted(got) == sorted(want), (name, sorted(set(got) ^ set(want)))
            for k, v in got.items():
                assert v == pytest.approx(want[k], rel=1e-9, abs=1e-12), (name, k, v, want[k])
            break
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        want = {key(s, i): ((int(x), int(y)), float(e), float(a)) for s, i, x, y, e, a in
                zip(z["st_s"][s0:s1], z["st_id"][s0:s1], z["st_x"][s0:s1], z["st_y"][s0:s1], z["st_e"][s0:s1], z["st_acc"][s0:s1])}
        pos, en, acc = env.agent_positions, env.agent_energies, env.agent_move_accumulator
        assert sorted(pos) == sorted(want), (name, t)
        assert all((pos[k], en[k], acc[k]) == want[k] for k in want), (name, t)
    env.close()
