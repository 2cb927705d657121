"""The dict adapters' HOST logic in the CPU suite: the same episode replays as tests/test_gpu_dict_adapters.py and
tests/test_gpu_parity.py::test_dict_adapter_replays_reference_episode, with the C oracle standing in for the CUDA library
(tests/oracle_batch.py).  What is tested here is the adapter code — reset tapes, row <-> dict mapping, dict orders, `agents`,
truncation rules, `training_metrics`, infos — against the reference's recordings; the CUDA path is the `-m gpu` suite."""
import pytest

import predpreygrass_b200.batched as batched
from tests import test_gpu_dict_adapters as G
from tests import test_gpu_parity as P
from tests.helpers import golden_cases
from tests.oracle_batch import OracleBatch


@pytest.fixture(autouse=True)
def _oracle_backend(monkeypatch):
    monkeypatch.setattr(batched, "BatchedPredPreyGrass", OracleBatch)


@pytest.mark.parametrize("name", ["base_default_s1", "base_default_s7_shuffle", "base_trunc_s2", "base_crowded_s1", "additive_crowded_s2",
                                  "kickback_crowded_s1", "dense_crowded_s3", "seasonal_default_s1"])
def test_base_adapter_over_the_oracle(name):
    P.test_dict_adapter_replays_reference_episode.__wrapped__(name) if hasattr(P.test_dict_adapter_replays_reference_episode, "__wrapped__") \
        else P.test_dict_adapter_replays_reference_episode(name)


@pytest.mark.parametrize("name", golden_cases(("eco",)))
def test_eco_adapter_over_the_oracle(name):
    G.test_eco_dict_adapter_replays_reference_episode(name)


@pytest.mark.parametrize("name", golden_cases(("stag",)))
def test_stag_adapter_over_the_oracle(name):
    G.test_stag_dict_adapter_replays_reference_episode(name)


@pytest.mark.parametrize("name", golden_cases(("mr", "inv", "coop")))
def test_trait_adapter_over_the_oracle(name):
    G.test_trait_dict_adapter_replays_reference_episode(name)


@pytest.mark.parametrize("name", golden_cases(("cad",)))
def test_cadence_adapter_over_the_oracle(name):
    G.test_cadence_dict_adapter_replays_reference_episode(name)


@pytest.mark.parametrize("name", ["coop_crowded_s3", "mr_default_s1"])
def test_trait_adapter_without_the_event_recorder(name):
    """`record_agent_events: False`: no per-step host replay — the exporters are empty, `get_all_agent_stats()` raises, and
    `training_metrics` leaves out exactly the keys that need the recorder (rank correlations, kinship)"""
    import json

    from predpreygrass_b200 import env_evolutionary as E
    from tests.helpers import load_golden

    z, cfg = load_golden(name)
    cls = {"mr": E.PredPreyGrassMetabolicRate, "coop": E.PredPreyGrassCooperation}[cfg.pop("variant")]
    cfg.update(cap_live=(250, 450), record_agent_events=False)
    env = cls(cfg)
    env.reset(seed=int(z["seed"]), options={"ppg_tape": (z["fallback_cells"], z["step_reals"])})
    names = ("predator", "prey")
    got = None
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        *_, term, trunc, infos = env.step({f"{names[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])})
        if term["__all__"] or trunc["__all__"]:
            got = infos["__all__"]["training_metrics"]
            break
    want = json.loads(str(z["metrics_json"]))
    missing = set(want) - set(got)
    assert missing and all("spearman" in k or "relatedness" in k for k in missing), sorted(missing)
    assert all(got[k] == pytest.approx(want[k], rel=1e-9, abs=1e-12) for k in got)
    assert env.per_step_agent_data == [] and env.agent_event_log == {}
    with pytest.raises(RuntimeError):
        env.get_all_agent_stats()
    env.close()
