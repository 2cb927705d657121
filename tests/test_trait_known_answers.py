"""The known-answer tests of the reference's own suites for the other heritable-trait variants, restated as scenarios on
the batched step and checked on BOTH sides (CPU oracle `-m "not gpu"`, CUDA path through the C-ABI `-m gpu`):
  MR-T   = predpreygrass/evolutionary/eco_evolutionary_metabolic_rate/tests/test_eco_evolutionary_validation.py
  INV-T  = predpreygrass/evolutionary/eco_evolutionary_investment/tests/test_eco_evolutionary_validation.py
  COOP-T = predpreygrass/evolutionary/eco_evolutionary_cooperation/tests/test_eco_evolutionary_validation.py
  CAD-T  = predpreygrass/evolutionary/eco_evolutionary_cadence/tests/test_eco_evolutionary_validation.py

Same method as tests/test_eco_known_answers.py: the reference tests teleport agents and overwrite energies / genomes,
then call one private method; here the same worlds are set up through the public inputs only (the replay tape gives
the founder counts, cells and founder trait values; the config gives the energies) and one whole `step()` runs.  The
expected numbers are the reference tests' own formulas, plus the basal decay of the whole step."""
import numpy as np
import pytest

from predpreygrass_b200.config import (CADENCE_CONFIG, COOPERATION_CONFIG, INVESTMENT_CONFIG, METABOLIC_CONFIG, VARIANT_ECO,
                                      make_config)
from tests.test_eco_known_answers import BACKENDS, TERM, TRUNC, World

STAY = 4            # (0, 0) of the 3x3 action table (MR:203-210)
MOVE_1_0 = 7        # (1, 0)
MOVE_0_1 = 5        # (0, 1)
G = METABOLIC_CONFIG["grid_size"]
FAR_GRASS = [G * (G - 1) + k for k in range(4)]
TINY = dict(n_possible_predators=8, n_possible_prey=8, initial_num_grass=4)  # MR-T:12-29
NO_BIRTHS = dict(predator_creation_energy_threshold=999.0, prey_creation_energy_threshold=999.0)
PRED, PREY = (0, 0), (1, 0)


def cell(x, y):
    return x * G + y


class TraitWorld(World):
    """one env of a trait variant on the oracle or on the GPU; `founders` = (n_pred, n_prey) of the episode"""

    def __init__(self, backend, base, overrides, founders, cells, traits, reals=()):
        cfgd = dict(base, **TINY)
        cfgd.update(n_initial_active_predators=founders[0], n_initial_active_prey=founders[1])
        cfgd.update(overrides)
        # MR-T:26-28 pins min == max so that the founder count is deterministic
        cfgd["n_initial_active_predators_min"], cfgd["n_initial_active_prey_min"] = founders
        self.cfg = make_config(cfgd, variant=VARIANT_ECO, cap_live=(32, 32), autoreset=False)
        self.backend = backend
        cells, traits = np.asarray(cells, np.int32), np.asarray(traits, np.float64)
        if backend == "oracle":
            from oracle.oracle import Oracle

            self.o = Oracle(self.cfg, 1)
            self.o.load_tape([np.zeros(0, np.int32)], [np.asarray(reals, np.float64)])
            self.out = self.o.env_reset_trait(0, founders[0], founders[1], cells, traits)
        else:
            from predpreygrass_b200.batched import BatchedPredPreyGrass

            self.g = BatchedPredPreyGrass(self.cfg, 1)
            ints = cells if self.cfg.trait_mode == 4 else np.concatenate([np.asarray(founders, np.int32), cells])  # cadence: fixed founder count
            self.g.load_tape([ints], [np.concatenate([traits, np.asarray(reals, np.float64)])])
            self.g.reset()
            self.out = self.g.outputs_numpy()

    def step(self, actions):
        if self.backend == "oracle":
            return super().step(actions)
        import torch

        out = self.out
        for s in range(2):
            a = np.full(max(1, out["n"][s]), STAY, np.int32)
            for r in range(out["n"][s]):
                if not out[f"flags{s}"][r] & 3:
                    a[r] = actions[(s, int(out[f"row_agent{s}"][r]))]
            self.g.actions[s][: len(a)].copy_(torch.from_numpy(a))
        self.g.step()
        self.out = self.g.outputs_numpy()
        return self.rows()


# ------------------------------------------------------------------------------------------------ metabolic rate
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("rate", [2.0, 0.5])
def test_metabolic_rate_scales_basal_cost_linearly(backend, rate):
    """MR-T:347-369 — basal cost 0.10 at rate 2.0 / 0.5: the prey loses 0.10 * rate per step (MR:555-561)"""
    w = TraitWorld(backend, METABOLIC_CONFIG, dict(NO_BIRTHS, basal_energy_cost_prey=0.10, initial_energy_prey=5.0, metabolic_rate_alpha=0.7),
                   (1, 1), [cell(5, 5), cell(15, 15)] + FAR_GRASS, [1.0, rate])
    w.step({PRED: STAY, PREY: STAY})
    assert w.state()[PREY]["energy"] == 5.0 - 0.10 * rate
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_metabolic_rate_scales_food_gain_sublinearly(backend):
    """MR-T:371-383 — a prey of rate 1.5 on a grass patch of energy 2.0 gains 2.0 * 1.5 ** 0.7 (MR:807); the basal cost of
    the step comes off first, the patch regrows first (capped at max_energy_grass = 2.0) and is emptied"""
    grass = [cell(15, 15)] + FAR_GRASS[:3]
    w = TraitWorld(backend, METABOLIC_CONFIG, dict(NO_BIRTHS, basal_energy_cost_prey=0.10, initial_energy_prey=5.0, metabolic_rate_alpha=0.7),
                   (1, 1), [cell(5, 5), cell(15, 15)] + grass, [1.0, 1.5])
    rows = w.step({PRED: STAY, PREY: STAY})
    assert w.state()[PREY]["energy"] == (5.0 - 0.10 * 1.5) + 2.0 * (1.5 ** 0.7)
    assert rows[PREY][2] & 0x10  # PPG_ROW_ATE
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_predator_gain_is_capped_and_scaled(backend):
    """MR:747-751 — bite = min(prey energy, max_energy_gain_per_prey); gain = bite * rate ** alpha; the prey is consumed"""
    ov = dict(NO_BIRTHS, n_initial_active_prey=2, max_energy_gain_per_prey=2.0, metabolic_rate_alpha=0.4, basal_energy_cost_predator=0.15,
              initial_energy_predator=5.0, initial_energy_prey=3.0, predator_satiation_cooldown=0)
    w = TraitWorld(backend, METABOLIC_CONFIG, ov, (1, 2), [cell(5, 5), cell(5, 5), cell(20, 20)] + FAR_GRASS, [1.25, 1.0, 1.0])
    rows = w.step({PRED: STAY, (1, 0): STAY, (1, 1): STAY})
    st = w.state()
    assert st[PRED]["energy"] == (5.0 - 0.15 * 1.25) + 2.0 * (1.25 ** 0.4)
    assert (1, 0) not in st and rows[(1, 0)][2] & TERM and not rows[(1, 0)][2] & TRUNC
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_satiation_cooldown_blocks_the_next_catch(backend):
    """MR:734-740,756-757 — a catch at current_step 0 with cooldown 3 sets satiation_until = 3: the predator standing on the
    second prey does not hunt while current_step < 3 (calls 2 and 3) and eats it in call 4"""
    ov = dict(NO_BIRTHS, n_initial_active_prey=3, predator_satiation_cooldown=3, basal_energy_cost_prey=0.0, initial_energy_prey=3.0,
              basal_energy_cost_predator=0.15, initial_energy_predator=5.0, max_energy_gain_per_prey=8.0)
    w = TraitWorld(backend, METABOLIC_CONFIG, ov, (1, 3), [cell(5, 5), cell(5, 5), cell(5, 6), cell(20, 20)] + FAR_GRASS, [1.0] * 4)
    prey = {(1, k): STAY for k in range(3)}
    w.step({**prey, PRED: STAY})               # call 1: eats prey_0
    prey.pop((1, 0))
    assert (1, 0) not in w.state()
    w.step({**prey, PRED: MOVE_0_1})           # call 2: steps onto prey_1, still digesting
    st = w.state()
    assert st[PRED]["xy"] == (5, 6) and (1, 1) in st
    w.step({**prey, PRED: STAY})               # call 3: current_step 2 < 3
    assert (1, 1) in w.state()
    rows = w.step({**prey, PRED: STAY})        # call 4: current_step 3, hunts again
    st = w.state()
    assert (1, 1) not in st and rows[(1, 1)][2] & TERM
    assert st[PRED]["energy"] == pytest.approx(5.0 - 4 * 0.15 + 3.0 + 3.0)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_mr_movement_cost_uses_distance_without_genome_multiplier(backend):
    """MR-T:669-692 — cost_per_cell * distance, whatever the metabolic rate (MR:531-539)"""
    ov = dict(NO_BIRTHS, basal_energy_cost_predator=0.2, movement_energy_cost_per_cell_predator=0.05, initial_energy_predator=10.0)
    w = TraitWorld(backend, METABOLIC_CONFIG, ov, (1, 1), [cell(10, 10), cell(20, 20)] + FAR_GRASS, [2.0, 1.0])
    w.step({PRED: MOVE_1_0, PREY: STAY})
    st = w.state()[PRED]
    assert st["xy"] == (11, 10) and st["energy"] == (10.0 - 0.2 * 2.0) - 0.05 * 1.0
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_density_cap_blocks_predator_births(backend):
    """MR:843-854 — predator_reproduction_max_ratio 0.5 with 1 predator and 1 prey: 1 >= 0.5 * 1, no birth although the
    energy is above the threshold; without the cap the same world has a birth"""
    for ratio, births in ((0.5, 0), (None, 1)):
        ov = dict(predator_creation_energy_threshold=10.0, prey_creation_energy_threshold=999.0, initial_energy_predator=20.0,
                  predator_reproduction_max_ratio=ratio, genome_mutation={"rate": 0.0, "std": 0.0})
        w = TraitWorld(backend, METABOLIC_CONFIG, ov, (1, 1), [cell(5, 5), cell(20, 20)] + FAR_GRASS, [1.0, 1.0])
        w.step({PRED: STAY, PREY: STAY})
        assert sum(1 for k in w.state() if k[0] == 0) == 1 + births, ratio
        w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_mr_offspring_receives_fixed_initial_energy_and_mutated_rate(backend):
    """MR-T:529-593 — the child starts with initial_energy_predator, the parent pays it, the child's rate is the parent's
    plus the mutation draw (rate 1.0: `u < rate` then `delta`)"""
    ov = dict(predator_creation_energy_threshold=10.0, prey_creation_energy_threshold=999.0, initial_energy_predator=20.25,
              basal_energy_cost_predator=0.25, genome_mutation={"rate": 1.0, "std": 0.01})
    w = TraitWorld(backend, METABOLIC_CONFIG, ov, (1, 1), [cell(5, 5), cell(20, 20)] + FAR_GRASS, [1.0, 1.0], reals=[0.5, 0.00390625])
    rows = w.step({PRED: STAY, PREY: STAY})
    st = w.state()
    assert st[(0, 1)]["energy"] == 20.25 and st[PRED]["energy"] == (20.25 - 0.25) - 20.25
    assert st[PRED]["speed"] == 1.0 and st[(0, 1)]["speed"] == 1.0 + 0.00390625
    assert rows[PRED][1] == 10.0 and (0, 1) in rows
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_time_limit_truncates_with_three_channel_rows(backend):
    """MR-T:240-275 — max_steps 1: everyone truncated, rows are (3, R, R) (the trait variants have no speed plane)"""
    w = TraitWorld(backend, METABOLIC_CONFIG, dict(NO_BIRTHS, max_steps=1), (1, 1), [cell(1, 1), cell(G - 2, G - 2)] + FAR_GRASS, [1.0, 1.0])
    rows = w.step({PRED: STAY, PREY: STAY})
    assert rows[PRED][0].shape == (3, 7, 7) and rows[PREY][0].shape == (3, 9, 9)
    for k in (PRED, PREY):
        assert rows[k][2] & TRUNC and not rows[k][2] & TERM
    assert w.env_flags() & 2 and not w.env_flags() & 1
    w.close()


# ------------------------------------------------------------------------------------------------ offspring investment
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("fraction", [0.60, 0.30])
def test_investment_fraction_determines_offspring_energy_and_parent_cost(backend, fraction):
    """INV-T:347-378 — child energy = parent energy * the PARENT's fraction, the parent pays exactly that (INV:546-557,890)"""
    ov = dict(predator_creation_energy_threshold=10.0, prey_creation_energy_threshold=999.0, initial_energy_predator_at_reset=20.25,
              energy_loss_per_step_predator=0.25, genome_mutation={"rate": 0.0, "std": 0.0})
    w = TraitWorld(backend, INVESTMENT_CONFIG, ov, (1, 1), [cell(5, 5), cell(20, 20)] + FAR_GRASS, [fraction, 0.35])
    w.step({PRED: STAY, PREY: STAY})
    st = w.state()
    parent = 20.25 - 0.25
    assert st[(0, 1)]["energy"] == parent * fraction
    assert st[PRED]["energy"] == parent - parent * fraction
    assert st[(0, 1)]["speed"] == fraction  # exact copy at mutation rate 0 (INV-T:299-314)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_investment_parent_cannot_reproduce_again_until_energy_refills(backend):
    """INV-T:458-476 — after one birth the parent is below the threshold: no second child in the next step"""
    ov = dict(predator_creation_energy_threshold=10.0, prey_creation_energy_threshold=999.0, initial_energy_predator_at_reset=12.0,
              energy_loss_per_step_predator=0.25, genome_mutation={"rate": 0.0, "std": 0.0})
    w = TraitWorld(backend, INVESTMENT_CONFIG, ov, (1, 1), [cell(5, 5), cell(20, 20)] + FAR_GRASS, [0.5, 0.35])
    w.step({PRED: STAY, PREY: STAY})
    assert sum(1 for k in w.state() if k[0] == 0) == 2
    w.step({PRED: STAY, (0, 1): STAY, PREY: STAY})
    assert sum(1 for k in w.state() if k[0] == 0) == 2
    w.close()


# ------------------------------------------------------------------------------------------------ cooperation
COOP_OV = dict(NO_BIRTHS, basal_energy_cost_predator=0.0, basal_energy_cost_prey=0.0, initial_energy_predator=3.0, initial_energy_prey=4.0,
               cooperation_range=2)


@pytest.mark.parametrize("backend", BACKENDS)
def test_donation_shares_gain_with_same_species_neighbor_in_range(backend):
    """COOP-T:648-675 — donor (rate 0.5) on (5, 5) eats a prey of energy 4.0; the other predator on (5, 6) receives
    0.5 * 4.0, the donor keeps the rest"""
    w = TraitWorld(backend, COOPERATION_CONFIG, COOP_OV, (2, 2), [cell(5, 5), cell(5, 6), cell(5, 5), cell(20, 20)] + FAR_GRASS,
                   [0.5, 0.0, 0.0, 0.0])
    w.step({(0, 0): STAY, (0, 1): STAY, (1, 0): STAY, (1, 1): STAY})
    st = w.state()
    assert st[(0, 1)]["energy"] == 3.0 + 0.5 * 4.0
    assert st[(0, 0)]["energy"] == 3.0 + (4.0 - 0.5 * 4.0)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_donation_is_split_equally_among_neighbours(backend):
    """COOP:560-573 — two predators in range: each receives rate * gain / 2"""
    w = TraitWorld(backend, COOPERATION_CONFIG, COOP_OV, (3, 2), [cell(5, 5), cell(5, 6), cell(7, 7), cell(5, 5), cell(20, 20)] + FAR_GRASS,
                   [0.75, 0.0, 0.0, 0.0, 0.0])
    w.step({(0, 0): STAY, (0, 1): STAY, (0, 2): STAY, (1, 0): STAY, (1, 1): STAY})
    st = w.state()
    assert st[(0, 1)]["energy"] == 3.0 + 0.75 * 4.0 / 2 and st[(0, 2)]["energy"] == 3.0 + 0.75 * 4.0 / 2
    assert st[(0, 0)]["energy"] == 3.0 + (4.0 - 0.75 * 4.0)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case", ["no_neighbour", "out_of_range", "zero_rate", "no_meal"])
def test_no_donation(backend, case):
    """COOP-T:678-765 — no eligible same-species neighbour / neighbour beyond the Chebyshev radius / rate 0 / nothing eaten:
    the other predator's energy does not change"""
    other = {"no_neighbour": None, "out_of_range": cell(5, 8), "zero_rate": cell(5, 6), "no_meal": cell(5, 6)}[case]
    rate = 0.0 if case == "zero_rate" else 0.8
    prey_cell = cell(15, 15) if case == "no_meal" else cell(5, 5)
    if other is None:
        w = TraitWorld(backend, COOPERATION_CONFIG, COOP_OV, (1, 2), [cell(5, 5), prey_cell, cell(20, 20)] + FAR_GRASS, [rate, 0.0, 0.0])
        w.step({(0, 0): STAY, (1, 0): STAY, (1, 1): STAY})
        assert w.state()[(0, 0)]["energy"] == 3.0 + 4.0
    else:
        w = TraitWorld(backend, COOPERATION_CONFIG, COOP_OV, (2, 2), [cell(5, 5), other, prey_cell, cell(20, 20)] + FAR_GRASS, [rate, 0.0, 0.0, 0.0])
        w.step({(0, 0): STAY, (0, 1): STAY, (1, 0): STAY, (1, 1): STAY})
        st = w.state()
        assert st[(0, 1)]["energy"] == 3.0
        assert st[(0, 0)]["energy"] == (3.0 if case == "no_meal" else 7.0)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_prey_share_grass_with_prey_only(backend):
    """COOP:828 — a prey of rate 0.5 eating grass (2.0) shares 1.0 with the prey next to it, nothing with the predator"""
    grass = [cell(15, 15)] + FAR_GRASS[:3]
    ov = dict(COOP_OV, initial_energy_prey=3.0)
    w = TraitWorld(backend, COOPERATION_CONFIG, ov, (1, 2), [cell(15, 16), cell(15, 15), cell(16, 16)] + grass, [0.0, 0.5, 0.0])
    w.step({(0, 0): STAY, (1, 0): STAY, (1, 1): STAY})
    st = w.state()
    assert st[(1, 0)]["energy"] == 3.0 + 1.0 and st[(1, 1)]["energy"] == 3.0 + 1.0 and st[(0, 0)]["energy"] == 3.0
    w.close()


# ------------------------------------------------------------------------------------------------ cadence
def cad_world(backend, overrides, speeds, accs, cells=None, founders=(1, 1), reals=()):
    """founder traits of a cadence world: the speeds, then the accumulator phases (CAD:1327)"""
    cells = [cell(10, 10), cell(20, 20)] + FAR_GRASS if cells is None else cells
    return TraitWorld(backend, CADENCE_CONFIG, dict(NO_BIRTHS, max_agent_age={"predator": None, "prey": None}, **overrides), founders, cells,
                      list(speeds) + list(accs), reals=reals)


def acc_of(w):
    a = w.o.read_env_acc(0) if w.backend == "oracle" else w.g.read_env_acc(0)
    return float(a[0][0]), float(a[1][0])


@pytest.mark.parametrize("backend", BACKENDS)
def test_slow_speed_is_frozen_when_accumulator_below_threshold(backend):
    """CAD-T:404-420 — speed 0.0: move rate 1 / max_cooldown = 0.1; from an empty accumulator the agent stays, the accumulator
    advances by the rate; the row of the next step carries the "stay only" action mask (0.1 + 0.1 < 1)"""
    w = cad_world(backend, {}, [0.0, 0.5], [0.0, 0.0])
    rows = w.step({PRED: MOVE_1_0, PREY: STAY})
    assert w.state()[PRED]["xy"] == (10, 10)
    assert acc_of(w)[0] == pytest.approx(0.1)
    assert rows[PRED][2] & 0x80  # PPG_ROW_FROZEN
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_fast_speed_moves_every_step(backend):
    """CAD-T:423-438 — speed 1.0: rate 1.0, the accumulator crosses 1.0 in one step: 0.0 + 1.0 - 1.0 = 0.0, never frozen"""
    w = cad_world(backend, {}, [1.0, 0.5], [0.0, 0.0])
    for t in range(3):
        rows = w.step({PRED: MOVE_1_0, PREY: STAY})
        assert w.state()[PRED]["xy"] == (11 + t, 10) and acc_of(w)[0] == 0.0
        assert not rows[PRED][2] & 0x80
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_accumulator_lets_a_mid_speed_agent_move_on_schedule(backend):
    """CAD:556-571,674-681 — speed 0.5, max_cooldown 5: rate 0.2 + 0.5 * 0.8 = 0.6; from 0.0 the accumulator reads 0.6 (frozen),
    then 1.2 -> moves -> 0.2, then 0.8 (frozen), then 1.4 -> moves -> 0.4"""
    w = cad_world(backend, dict(max_cooldown=5), [0.5, 0.5], [0.0, 0.0])
    xs, accs = [], []
    for _ in range(4):
        w.step({PRED: MOVE_1_0, PREY: STAY})
        xs.append(w.state()[PRED]["xy"][0]); accs.append(acc_of(w)[0])
    assert xs == [10, 11, 11, 12]
    rate = 1.0 / 5 + 0.5 * (1.0 - 1.0 / 5)
    a, want = 0.0, []
    for _ in range(4):
        a = a + rate
        a = a - 1.0 if a >= 1.0 else a
        want.append(a)
    assert accs == want
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("speed,action,cost", [(1.0, STAY, 0.0), (0.5, MOVE_1_0, 0.05 * 1.0 * (0.5 ** 2))])
def test_cadence_costs(backend, speed, action, cost):
    """CAD-T:441-468 stationary at speed 1.0: only the basal 0.2 * (1 + 1.0 * 1.0); CAD-T:471-500 one cell at speed 0.5:
    basal 0.2 * (1 + 0.5) plus 0.05 * 1 * 0.5 ** 2 (the accumulator is pre-loaded to 1.0 so that the move executes)"""
    ov = dict(energy_loss_per_step_predator=0.2, movement_energy_cost_per_cell_predator=0.05, movement_speed_cost_exponent=2.0,
              metabolic_speed_coeff=1.0, initial_energy_predator=10.0)
    w = cad_world(backend, ov, [speed, 0.5], [1.0, 0.0])
    w.step({PRED: action, PREY: STAY})
    assert w.state()[PRED]["energy"] == pytest.approx(10.0 - 0.2 * (1.0 + 1.0 * speed) - cost)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_cadence_catch_radius_and_nearest_prey(backend):
    """CAD:837-848 — the predator catches within Chebyshev distance 1: of a prey on a diagonal neighbour and one on its own
    cell the nearer one; the prey is eaten whole (CAD:857), the other one lives on"""
    ov = dict(energy_loss_per_step_predator=0.0, energy_loss_per_step_prey=0.0, metabolic_speed_coeff=0.0, initial_energy_predator=5.0,
              initial_energy_prey=3.0)
    cells = [cell(10, 10), cell(11, 11), cell(10, 10), cell(20, 20)] + FAR_GRASS
    w = cad_world(backend, ov, [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], cells=cells, founders=(1, 3))
    rows = w.step({PRED: STAY, (1, 0): STAY, (1, 1): STAY, (1, 2): STAY})
    st = w.state()
    assert (1, 1) not in st and rows[(1, 1)][2] & TERM      # the prey on the predator's cell (distance 0) goes first
    assert (1, 0) in st and st[PRED]["energy"] == 5.0 + 3.0
    rows = w.step({PRED: STAY, (1, 0): STAY, (1, 2): STAY})
    st = w.state()
    assert (1, 0) not in st and rows[(1, 0)][2] & TERM      # now the diagonal neighbour
    assert st[PRED]["energy"] == 5.0 + 3.0 + 3.0 and (1, 2) in st
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_cadence_speed_plane_is_the_raw_genome_value(backend):
    """CAD:746 — `spatial[n_ch] = float(genome.speed)`, not normalised by the trait bounds"""
    w = cad_world(backend, dict(trait_bounds={"speed": (0.0, 2.0)}), [0.75, 0.25], [0.0, 0.0])
    rows = w.rows()
    assert rows[PRED][0].shape == (4, 7, 7) and np.all(rows[PRED][0][3] == np.float32(0.75))
    assert np.all(rows[PREY][0][3] == np.float32(0.25))
    w.close()


def test_trait_config_mandatory_keys_raise():
    """the trait variants read their mandatory keys with `config[...]` (MR:38-72): a missing one raises KeyError"""
    for base in (METABOLIC_CONFIG, INVESTMENT_CONFIG, COOPERATION_CONFIG):
        bad = dict(base)
        bad.pop("max_energy_grass")
        with pytest.raises(KeyError):
            make_config(bad, variant=VARIANT_ECO)
    with pytest.raises(ValueError):
        make_config(dict(METABOLIC_CONFIG, genome_neutral_drift_control=True), variant=VARIANT_ECO)
