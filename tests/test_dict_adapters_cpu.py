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
