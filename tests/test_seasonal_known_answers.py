"""The reference's own tests of base_environment_seasonal
(predpreygrass/non_evolutionary/base_environment_seasonal/tests/test_seasonal_grass_regrowth.py, cited as REF:lines)
restated on the batched step: CPU oracle (`-m "not gpu"`) and CUDA path through the C-ABI (`-m gpu`).  BASE itself
has no reference tests; this suite is the one place where the reference pins numbers of the BASE step (the grass
regrowth of BASE:252-256 with the seasonal gain of SEASON:268-271)."""
import numpy as np
import pytest

from predpreygrass_b200.config import BASE_CONFIG, SEASONAL_CONFIG, make_config, season_multiplier

BACKENDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]
STAY, DOWN = 4, 7  # BASE:96-106: a -> (a // 3 - 1, a % 3 - 1); 7 = (1, 0)
G = BASE_CONFIG["grid_size"]


def test_season_multiplier_phase_boundaries():
    """REF:15-36 — steps 0-4 high, 5-9 low, 10-14 high again"""
    cfg = dict(SEASONAL_CONFIG, season_length_steps=5, season_high_multiplier=1.5, season_low_multiplier=0.5)
    assert [season_multiplier(cfg, t) for t in (0, 1, 4)] == [1.5] * 3
    assert [season_multiplier(cfg, t) for t in (5, 6, 9)] == [0.5] * 3
    assert [season_multiplier(cfg, t) for t in (10, 14)] == [1.5] * 2
    c = make_config(cfg)
    assert c.season_length_steps == 5 and list(c.season_multiplier) == [1.5, 0.5]
    assert make_config(BASE_CONFIG).season_length_steps == 0  # BASE: no seasons


def _world(backend, cfg, cells):
    c = make_config(cfg, cap_live=(32, 32), autoreset=False)
    cells = np.asarray(cells, np.int32)
    if backend == "oracle":
        from oracle.oracle import Oracle

        o = Oracle(c, 1)
        o.env_reset_cells(0, cells)

        def step(acts):
            keys = list(acts)
            o.env_step_ordered(0, [k[0] for k in keys], [k[1] for k in keys], [acts[k] for k in keys])
            return o.read_env(0)
        return step, o.close
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    g = BatchedPredPreyGrass(c, 1)
    g.load_tape([cells])
    g.reset()

    def step(acts):
        out = g.outputs_numpy()
        for s in range(2):
            a = np.full(max(1, out["n"][s]), STAY, np.int32)
            for r in range(out["n"][s]):
                a[r] = acts[(s, int(out[f"row_agent{s}"][r]))]
            g.actions[s][: len(a)].copy_(torch.from_numpy(a))
        g.step()
        return g.read_env(0)
    return step, g.close


@pytest.mark.parametrize("backend", BACKENDS)
def test_season_disabled_reproduces_flat_baseline(backend):
    """REF:39-47 — both multipliers 1.0: the mechanism is a no-op whatever the season length"""
    cells = [0, G * G - 1, 5 * G + 5, 9 * G + 9]  # predator, prey, two grass patches; the prey sits far from everything
    flat, close_a = _world(backend, dict(BASE_CONFIG, n_initial_active_predator=1, n_initial_active_prey=1, initial_num_grass=2,
                                         initial_energy_grass=2.0), cells)
    seas, close_b = _world(backend, dict(BASE_CONFIG, n_initial_active_predator=1, n_initial_active_prey=1, initial_num_grass=2,
                                         season_length_steps=3, season_high_multiplier=1.0, season_low_multiplier=1.0), cells)
    for t in range(20):
        acts = {(0, 0): (t * 5) % 9, (1, 0): (t * 7 + 3) % 9}
        a, b = flat(acts), seas(acts)
        assert np.array_equal(a["grass_energy"], b["grass_energy"]) and np.array_equal(a["energy"][1], b["energy"][1])
        assert np.array_equal(a["xy"][0], b["xy"][0]) and np.array_equal(a["xy"][1], b["xy"][1])
    close_a(); close_b()


@pytest.mark.parametrize("backend", BACKENDS)
def test_grass_regrows_faster_in_abundant_phase_than_scarce_phase(backend):
    """REF:54-84 — season length 3, multipliers 1.5 / 0.5: a patch at 0 gains 3 * gain * 1.5 over an abundant phase and
    3 * gain * 0.5 over a scarce one.  The reference test zeroes a patch by hand; here a prey standing on it eats it in
    step 0 (regrowth precedes eating, BASE:252-256 before :347-380) and walks away in step 1."""
    gain = BASE_CONFIG["energy_gain_per_step_grass"]
    cfg = dict(BASE_CONFIG, n_initial_active_predator=1, n_initial_active_prey=1, initial_num_grass=2,
               season_length_steps=3, season_high_multiplier=1.5, season_low_multiplier=0.5,
               prey_creation_energy_threshold=999.0, predator_creation_energy_threshold=999.0)
    step, close = _world(backend, cfg, [0, 10 * G + 10, 10 * G + 10, 20 * G + 20])  # the prey stands on grass patch 0
    e = {}
    for t in range(9):
        st = step({(0, 0): STAY, (1, 0): DOWN if t == 1 else STAY})
        e[t] = float(st["grass_energy"][0])
    assert e[0] == 0.0                                     # eaten in step 0
    assert e[2] == 0.0 + gain * 1.5 + gain * 1.5           # steps 1, 2: abundant (bit-exact: the same fp64 sums)
    scarce = e[5] - e[2]                                   # steps 3, 4, 5
    abundant = e[8] - e[5]                                 # steps 6, 7, 8
    assert scarce == pytest.approx(3 * gain * 0.5) and abundant == pytest.approx(3 * gain * 1.5)
    assert abundant > scarce
    close()
