"""Host side of the ECO / STAG dict adapters (`-m "not gpu"`): `reset(seed)` must hand the device exactly the draws
the reference's own reset makes from `np.random.default_rng(seed)` (ECO:130,208-215,1752; STAG:273,2140-2169).
Pinned on the recordings of the unmodified reference classes in tests/golden/ (make_golden_eco.py /
make_golden_stag.py store the seed next to the placement, founder speeds, facings and raw traits they observed)."""
import numpy as np
import pytest

from predpreygrass_b200.env_evolutionary import reference_reset_tape_eco, reference_reset_tape_stag
from tests.helpers import golden_cases, load_golden


@pytest.mark.parametrize("name", golden_cases(("eco",)))
def test_eco_reset_tape_equals_the_reference_reset(name):
    z, cfg = load_golden(name)
    cells, speeds = reference_reset_tape_eco(int(z["seed"]), cfg)
    assert np.array_equal(cells, z["init_cells"])
    assert np.array_equal(speeds, z["founder_speed"])  # bit-exact float64, after the clip to trait_bounds


@pytest.mark.parametrize("name", golden_cases(("stag",)))
def test_stag_reset_tape_equals_the_reference_reset(name):
    z, cfg = load_golden(name)
    cells, facing, traits = reference_reset_tape_stag(int(z["seed"]), cfg)
    assert np.array_equal(cells, z["init_cells"])
    assert np.array_equal(facing, z["founder_facing"])
    assert np.array_equal(traits, z["founder_trait_raw"])


def test_adapters_need_a_config_like_the_reference():
    from predpreygrass_b200.env_evolutionary import PredPreyGrassEco, PredPreyGrassStag

    for cls in (PredPreyGrassEco, PredPreyGrassStag):
        with pytest.raises(ValueError):  # ECO:21-22, STAG:21-22
            cls(None)


TRAIT_OF = {"mr": "metabolic_rate", "inv": "offspring_investment_fraction", "coop": "cooperation_rate"}


@pytest.mark.parametrize("name", golden_cases(("mr", "inv", "coop")))
def test_trait_reset_tape_equals_the_reference_reset(name):
    """founder counts (MR:189-192), founder trait values (genome.py founder_genome) and cells (MR:1473) of the variants"""
    from predpreygrass_b200.env_evolutionary import reference_reset_tape_trait

    z, cfg = load_golden(name)
    ints, values = reference_reset_tape_trait(int(z["seed"]), cfg, TRAIT_OF[cfg["variant"]])
    assert np.array_equal(ints[:2], z["n_found"])
    assert np.array_equal(ints[2:], z["init_cells"])
    assert np.array_equal(values, z["founder_trait"])


@pytest.mark.parametrize("name", golden_cases(("cad",)))
def test_cadence_reset_tape_equals_the_reference_reset(name):
    """per founder a speed draw then an accumulator phase (CAD:1324-1327), then the cells"""
    from predpreygrass_b200.env_evolutionary import reference_reset_tape_cadence

    z, cfg = load_golden(name)
    cells, reals = reference_reset_tape_cadence(int(z["seed"]), cfg)
    assert np.array_equal(cells, z["init_cells"])
    assert np.array_equal(reals, np.concatenate([z["founder_trait"], z["founder_acc"]]))
