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
