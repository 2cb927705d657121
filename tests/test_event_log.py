"""`per_step_agent_data` (ECO:426-446) and `agent_event_log` (ECO:1488-1500, ...) of the ECO dict adapter against recordings
of the unmodified reference class (tests/golden/eco_events_*.json.gz, made by tests/golden/make_golden_eco_events.py).

The adapter's host logic is the same on both backends: the CPU test drives it over the C oracle (tests/oracle_batch.py), the
GPU test over the CUDA library.  Everything is compared for equality — positions, float64 energies, the four energy
deltas, ages, parents, offspring lists, every event with its step, ids, bite size and energy."""
import glob
import gzip
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN_DIR, load_golden

CASES = sorted(os.path.basename(p)[len("eco_events_"):-len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "eco_events_*.json.gz")))


def _plain(obj):
    if isinstance(obj, dict):
        return {str(k): _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


def _replay(case):
    from predpreygrass_b200.env_evolutionary import PredPreyGrassEco

    z, cfg = load_golden("eco_" + case)
    cfg.pop("variant")
    cfg["cap_live"] = (min(cfg["n_possible_predators"], 250), min(cfg["n_possible_prey"], 250))
    want = json.loads(gzip.open(os.path.join(GOLDEN_DIR, f"eco_events_{case}.json.gz")).read())
    env = PredPreyGrassEco(cfg)
    env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    names = ("predator", "prey")
    for t in range(want["steps"]):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        acts = {f"{names[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
        env.step(acts)
    got_steps, got_log = _plain(env.per_step_agent_data), _plain(env.agent_event_log)
    assert env._events.inexact_chains == 0
    assert len(got_steps) == len(want["per_step_agent_data"]) == want["steps"]
    for t, (g, w) in enumerate(zip(got_steps, want["per_step_agent_data"])):
        assert sorted(g) == sorted(w), (case, t)  # the recording is a JSON object with sorted keys: the order is not in it
        for a in w:
            assert g[a] == w[a], (case, t, a, g[a], w[a])
    assert sorted(got_log) == sorted(want["agent_event_log"]), case
    for a, w in want["agent_event_log"].items():
        for k in w:
            assert got_log[a][k] == w[k], (case, a, k, got_log[a][k], w[k])
    # get_all_agent_stats() (ECO:1676-1682): every record, every field, in the reference's order
    got_stats = _plain(env.get_all_agent_stats())
    assert list(got_stats) == list(want["agent_stats_order"]) if "agent_stats_order" in want else sorted(got_stats) == sorted(want["agent_stats"])
    for a, w in want["agent_stats"].items():
        assert sorted(got_stats[a]) == sorted(w), (case, a, sorted(set(got_stats[a]) ^ set(w)))
        for k in w:
            assert got_stats[a][k] == w[k], (case, a, k, got_stats[a][k], w[k])
    assert env._events.ambiguous_final_moves == 0
    assert env.get_total_offspring_by_type() == want["offspring_by_type"]
    env.close()


@pytest.mark.parametrize("case", CASES)
def test_event_log_host_logic_over_the_oracle(case, monkeypatch):
    import predpreygrass_b200.batched as batched
    from tests.oracle_batch import OracleBatch

    monkeypatch.setattr(batched, "BatchedPredPreyGrass", OracleBatch)
    _replay(case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_event_log_on_the_device(case):
    _replay(case)


@pytest.mark.gpu
def test_snapshot_restore_carries_the_exporters():
    """get_state_snapshot / restore_state_snapshot (ECO:1277-1376) of the dict adapter: the steps after a restore reproduce the
    steps after the snapshot, exporters included"""
    from predpreygrass_b200.env_evolutionary import PredPreyGrassEco

    z, cfg = load_golden("eco_default_s1")
    cfg.pop("variant")
    cfg["cap_live"] = (min(cfg["n_possible_predators"], 250), min(cfg["n_possible_prey"], 250))
    env = PredPreyGrassEco(cfg)
    env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    names = ("predator", "prey")

    def acts(t):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        return {f"{names[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}

    for t in range(20):
        env.step(acts(t))
    snap = env.get_state_snapshot()
    first = [env.step(acts(t)) for t in range(20, 30)]
    data1, log1 = _plain(env.per_step_agent_data), _plain(env.agent_event_log)
    env.restore_state_snapshot(snap)
    assert len(env.per_step_agent_data) == 20
    second = [env.step(acts(t)) for t in range(20, 30)]
    for (o1, r1, t1, u1, _), (o2, r2, t2, u2, _) in zip(first, second):
        assert list(o1) == list(o2) and r1 == r2 and t1 == t2 and u1 == u2
        assert all(np.array_equal(o1[k], o2[k]) for k in o1)
    assert _plain(env.per_step_agent_data) == data1 and _plain(env.agent_event_log) == log1
    assert env._events.inexact_chains == 0
    env.close()


TRAIT_CASES = sorted(os.path.basename(p)[len("trait_events_"):-len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "trait_events_*.json.gz")))


def _replay_trait(case):
    """the exporters of PredPreyGrassMetabolicRate / Investment / Cooperation / Cadence against recordings of the unmodified reference classes
    (tests/golden/trait_events_*.json.gz, made by tests/golden/make_golden_trait_events.py), and the order of the agent records"""
    from predpreygrass_b200 import env_evolutionary as E

    z, cfg = load_golden(case)
    cls = {"mr": E.PredPreyGrassMetabolicRate, "inv": E.PredPreyGrassInvestment, "coop": E.PredPreyGrassCooperation,
           "cad": E.PredPreyGrassCadence}[cfg.pop("variant")]
    cfg["cap_live"] = (min(cfg["n_possible_predators"], 250), min(cfg["n_possible_prey"], 450))
    if cls is E.PredPreyGrassCadence:
        cfg["record_step_data"] = True  # CAD:83,422: per_step_agent_data is optional there
    want = json.loads(gzip.open(os.path.join(GOLDEN_DIR, f"trait_events_{case}.json.gz")).read())
    env = cls(cfg)
    env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    names = ("predator", "prey")
    infos = {}
    for t in range(want["steps"]):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        acts = {f"{names[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
        *_, infos = env.step(acts)
    got_steps, got_log = _plain(env.per_step_agent_data), _plain(env.agent_event_log)
    assert env._events.inexact_chains == 0
    assert len(got_steps) == len(want["per_step_agent_data"]) == want["steps"]
    for t, (g, w) in enumerate(zip(got_steps, want["per_step_agent_data"])):
        assert sorted(g) == sorted(w), (case, t)
        for a in w:
            assert g[a] == w[a], (case, t, a, g[a], w[a])
    assert sorted(got_log) == sorted(want["agent_event_log"]), case
    for a, w in want["agent_event_log"].items():
        for k in w:
            assert got_log[a][k] == w[k], (case, a, k, got_log[a][k], w[k])
    assert env._events.record_order() == want["record_order"], case
    # get_all_agent_stats() (MR:1406-1412): every record, every field, in the reference's order
    got_stats = _plain(env.get_all_agent_stats())
    assert list(got_stats) == want["record_order"]
    for a, w in want["agent_stats"].items():
        assert sorted(got_stats[a]) == sorted(w), (case, a, sorted(set(got_stats[a]) ^ set(w)))
        for k in w:
            assert got_stats[a][k] == w[k], (case, a, k, got_stats[a][k], w[k])
    assert env.get_total_offspring_by_type() == want["offspring_by_type"]
    if not want["ended"]:
        assert env.get_total_energy_by_type() == want["energy_by_type"]
    if want["ended"]:
        got = infos["__all__"]["training_metrics"]
        for k, v in want["spearman"].items():
            assert got[k] == v, (case, k, got[k], v)
    env.close()


@pytest.mark.parametrize("case", TRAIT_CASES)
def test_trait_event_log_host_logic_over_the_oracle(case, monkeypatch):
    import predpreygrass_b200.batched as batched
    from tests.oracle_batch import OracleBatch

    monkeypatch.setattr(batched, "BatchedPredPreyGrass", OracleBatch)
    _replay_trait(case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", TRAIT_CASES)
def test_trait_event_log_on_the_device(case):
    _replay_trait(case)


_CROWDED = dict(grid_size=9, initial_num_grass=24, n_initial_active_predators=6, n_initial_active_prey=14,
                predator_creation_energy_threshold=6.0, prey_creation_energy_threshold=4.5, energy_gain_per_step_grass=0.3,
                predator_obs_range=5, prey_obs_range=7, n_possible_predators=300, n_possible_prey=400, max_steps=50)
_RICH = dict(energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0, predator_creation_energy_threshold=8.0, max_steps=60)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("fam", ["eco", "mr", "inv", "coop", "cad"])
def test_recorders_agree_with_the_backend_on_random_episodes(fam, seed, monkeypatch):
    """Beyond the recordings: on seeded random episodes (crowded 9 x 9 worlds with shuffled action dicts, reproduction-heavy
    worlds, carcasses / age caps / juvenile predators for ECO) the exporters' host replay must land on the backend's state every
    step — who is gone, the survivors' cells and float64 energies, the eating / reproduction flags (`inexact_chains` == 0)."""
    import predpreygrass_b200.batched as batched
    from predpreygrass_b200 import config as Cfg
    from predpreygrass_b200 import env_evolutionary as E
    from tests.oracle_batch import OracleBatch

    monkeypatch.setattr(batched, "BatchedPredPreyGrass", OracleBatch)
    cls, base = {"eco": (E.PredPreyGrassEco, Cfg.ECO_CONFIG), "mr": (E.PredPreyGrassMetabolicRate, Cfg.METABOLIC_CONFIG),
                 "inv": (E.PredPreyGrassInvestment, Cfg.INVESTMENT_CONFIG), "coop": (E.PredPreyGrassCooperation, Cfg.COOPERATION_CONFIG),
                 "cad": (E.PredPreyGrassCadence, Cfg.CADENCE_CONFIG)}[fam]
    over = dict(_CROWDED if seed % 2 else _RICH)
    if fam == "eco":
        over.update(max_energy_gain_per_prey=[1.0, 2.5, float("inf")][seed % 3], max_agent_age={"predator": [None, 30][seed % 2], "prey": 25},
                    carcass_only_predator_age={"predator": [None, 5][seed % 2]})
    cfg = dict(base, **over, cap_live=(250, 450), record_step_data=True)
    env = cls(cfg)
    rng = np.random.default_rng(100 + seed)
    env.reset(seed=seed)
    steps = 0
    for _ in range(over["max_steps"]):
        keys = list(env.agents)
        if seed % 2:
            rng.shuffle(keys)
        *_, term, trunc, _ = env.step({a: int(rng.integers(env.action_spaces[a].n)) for a in keys})
        steps += 1
        if term["__all__"] or trunc["__all__"]:
            break
    assert steps >= 2 and env._events.inexact_chains == 0, (fam, seed, steps, env._events.inexact_chains)
    assert len(env.per_step_agent_data) == steps and len(env.agent_event_log) >= len(env.per_step_agent_data[0])
    env.close()
