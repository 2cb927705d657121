"""Config keys the device does not implement are rejected loudly, never ignored (`-m "not gpu"`)."""
import pytest

from predpreygrass_b200.config import ECO_CONFIG, STAG_CONFIG, VARIANT_ECO, VARIANT_STAG, make_config


def test_stag_walls_and_line_of_sight_keys_are_read():
    c = make_config(dict(STAG_CONFIG, manual_wall_positions=[(3, 3), (3, 3), (40, 1), (0, 29)], respect_los_for_movement=True,
                         include_visibility_channel=True, mask_observation_with_visibility=True), variant=VARIANT_STAG)
    assert c.n_walls == 2 and list(c._walls) == [29, 93]  # duplicates and out-of-bounds cells dropped (STAG:2112-2122)
    assert c.respect_los_for_movement == 1 and c.include_visibility_channel == 1
    c = make_config(STAG_CONFIG, variant=VARIANT_STAG)
    assert c.grid_size == 30 and c.n_walls == 0 and c.include_visibility_channel == 0


def test_eco_lineage_coefficients_are_read_per_role():
    c = make_config(dict(ECO_CONFIG, lineage_reward_coeff={"predator": 0.5, "prey": 0.0}), variant=VARIANT_ECO)
    assert list(c.lineage_reward_coeff) == [0.5, 0.0]
    assert list(make_config(dict(ECO_CONFIG, lineage_reward_coeff=0.25), variant=VARIANT_ECO).lineage_reward_coeff) == [0.25, 0.25]
    assert make_config(ECO_CONFIG, variant=VARIANT_ECO).action_range == 5


def test_eco_mandatory_keys_raise_like_the_reference():
    cfg = dict(ECO_CONFIG)
    del cfg["max_energy_grass"]  # ECO:83 reads config["max_energy_grass"]
    with pytest.raises(KeyError):
        make_config(cfg, variant=VARIANT_ECO)
