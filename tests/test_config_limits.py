"""Config keys the device does not implement are rejected loudly, never ignored (`-m "not gpu"`)."""
import pytest

from predpreygrass_b200.config import ECO_CONFIG, STAG_CONFIG, VARIANT_ECO, VARIANT_STAG, make_config


def test_stag_walls_and_line_of_sight_are_rejected():
    with pytest.raises(ValueError):
        make_config(dict(STAG_CONFIG, manual_wall_positions=[(3, 3)]), variant=VARIANT_STAG)
    for k in ("mask_observation_with_visibility", "include_visibility_channel", "respect_los_for_movement"):
        with pytest.raises(ValueError):
            make_config(dict(STAG_CONFIG, **{k: True}), variant=VARIANT_STAG)
    assert make_config(STAG_CONFIG, variant=VARIANT_STAG).grid_size == 30  # the BASELINE config itself is fine


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
