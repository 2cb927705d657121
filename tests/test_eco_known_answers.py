"""The known-answer tests of the reference's own ECO suite
(predpreygrass/evolutionary/eco_evolutionary/tests/test_eco_evolutionary_validation.py, cited per test as REF:lines),
restated as scenarios on the batched step and checked on BOTH sides: the CPU oracle (`-m "not gpu"`) and the CUDA
path through the C-ABI (`-m gpu`).

The reference tests build a tiny env (1 predator, 1 prey, 4 grass, REF:12-26), teleport agents by editing the Python
dicts (REF:29-40), overwrite energies / genomes and call one private method or one `step()`.  Here the same worlds are
set up through the public inputs only: the replay tape chooses the cells and the founders' speeds
(include/ppg.h ppg_tape), the config chooses the energies, and one whole `step()` runs.  Expected numbers are the
reference tests' own formulas."""
import numpy as np
import pytest

from predpreygrass_b200.config import ECO_CONFIG, VARIANT_ECO, make_config

STAY = 12           # (0, 0) in the 5x5 action table (ECO:225-232: a -> (a // 5 - 2, a % 5 - 2))
MOVE_2_0 = 22       # (2, 0)
TINY = dict(n_initial_active_predators=1, n_initial_active_prey=1, n_possible_predators=8, n_possible_prey=8, initial_num_grass=4)  # REF:12-26
G = ECO_CONFIG["grid_size"]
FAR_GRASS = [G * (G - 1) + k for k in range(4)]  # bottom row, away from every scenario


def cell(x, y):
    return x * G + y


class World:
    """one env on the oracle or on the GPU, same calls"""

    def __init__(self, backend, overrides, cells, speeds, reals=()):
        cfgd = dict(ECO_CONFIG, **TINY)
        cfgd.update(overrides)
        self.cfg = make_config(cfgd, variant=VARIANT_ECO, cap_live=(32, 32), autoreset=False)
        self.backend = backend
        cells, speeds = np.asarray(cells, np.int32), np.asarray(speeds, np.float64)
        if backend == "oracle":
            from oracle.oracle import Oracle

            self.o = Oracle(self.cfg, 1)
            self.o.load_tape([np.zeros(0, np.int32)], [np.asarray(reals, np.float64)])
            self.out = self.o.env_reset_eco(0, cells, speeds)
        else:
            from predpreygrass_b200.batched import BatchedPredPreyGrass

            self.g = BatchedPredPreyGrass(self.cfg, 1)
            self.g.load_tape([cells], [np.concatenate([speeds, np.asarray(reals, np.float64)])])
            self.g.reset()
            self.out = self.g.outputs_numpy()

    def rows(self):
        """{(species, id): (obs, reward, flags)} of the last output"""
        out, res = self.out, {}
        for s in range(2):
            n = out["n"][s] if "n" in out else out["n_old"][s] + out["n_new"][s]
            for r in range(n):
                res[(s, int(out[f"row_agent{s}"][r]))] = (out[f"obs{s}"][r], float(out[f"reward{s}"][r]), int(out[f"flags{s}"][r]))
        return res

    def step(self, actions):
        """actions: {(species, id): action} for every live agent"""
        if self.backend == "oracle":
            keys = list(actions)
            self.out = self.o.env_step_ordered(0, np.array([k[0] for k in keys], np.int8), np.array([k[1] for k in keys], np.int32),
                                               np.array([actions[k] for k in keys], np.int8))
        else:
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

    def state(self):
        st = self.o.read_env_eco(0) if self.backend == "oracle" else self.g.read_env_eco(0)
        res = {}
        for s in range(2):
            for k, i in enumerate(st["ids"][s]):
                res[(s, int(i))] = dict(xy=tuple(int(v) for v in st["xy"][s][k]), energy=float(st["energy"][s][k]),
                                        age=int(st["age"][s][k]), speed=float(st["speed"][s][k]))
        return res

    def env_flags(self):
        return int(self.out["env_flags"][0])

    def close(self):
        (self.o if self.backend == "oracle" else self.g).close()


BACKENDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]
PRED, PREY = (0, 0), (1, 0)
NO_BIRTHS = dict(predator_creation_energy_threshold=999.0, prey_creation_energy_threshold=999.0)
TERM, TRUNC = 1, 2


@pytest.mark.parametrize("backend", BACKENDS)
def test_every_acted_agent_gets_next_or_final_observation(backend):
    """REF:144-174 — 4 predators, 6 prey, 30 grass on 12x12, ten random steps: whoever acted is in the next output"""
    rng = np.random.default_rng(456)
    ov = dict(max_steps=120, grid_size=12, n_initial_active_predators=4, n_initial_active_prey=6, n_possible_predators=80,
              n_possible_prey=160, initial_num_grass=30)
    cells = rng.choice(144, size=40, replace=False)
    w = World(backend, ov, cells, np.ones(10))
    rows = w.rows()
    for _ in range(10):
        acts = {k: int(rng.integers(25)) for k, v in rows.items() if not v[2] & (TERM | TRUNC)}
        rows = w.step(acts)
        for k in acts:
            assert k in rows and rows[k][0].shape[0] == 4  # next or final observation, 3 grid channels + speed plane
        if w.env_flags() & 3:
            break
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_agent_emits_max_age_termination(backend):
    """REF:205-222 — max_agent_age predator 4: the step that takes the age to the limit terminates the agent"""
    w = World(backend, dict(NO_BIRTHS, max_agent_age={"predator": 4, "prey": 400}), [cell(5, 5), cell(20, 20)] + FAR_GRASS, [1.0, 1.0])
    for t in range(1, 5):
        rows = w.step({PRED: STAY, PREY: STAY})
        assert bool(rows[PRED][2] & TERM) == (t == 4), t
    assert PRED not in w.state() or w.env_flags() & 1
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_terminal_reward_on_predation_and_extinction(backend):
    """REF:236-265 — predator and prey on (5, 5), both stay: the prey is caught (reward = penalty_prey_caught = 0,
    terminated, not truncated, still observed), the episode ends by extinction so the predator terminates as well"""
    w = World(backend, NO_BIRTHS, [cell(5, 5), cell(5, 5)] + FAR_GRASS, [1.0, 1.0])
    rows = w.step({PRED: STAY, PREY: STAY})
    assert PREY in rows and PRED in rows
    assert rows[PREY][1] == 0.0 and rows[PREY][2] & TERM and not rows[PREY][2] & TRUNC
    assert rows[PRED][2] & TERM and not rows[PRED][2] & TRUNC
    assert w.env_flags() & 1 and not w.env_flags() & 2  # terminations["__all__"], not truncations["__all__"]
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_eaten_prey_is_not_returned_as_next_actor(backend):
    """REF:268-293 — two prey, one under the predator: it gets a final row, the other one lives on"""
    ov = dict(NO_BIRTHS, n_initial_active_prey=2)
    w = World(backend, ov, [cell(5, 5), cell(5, 5), cell(G - 2, G - 2)] + FAR_GRASS, [1.0, 1.0, 1.0])
    rows = w.step({PRED: STAY, (1, 0): STAY, (1, 1): STAY})
    assert rows[(1, 0)][2] & TERM and not rows[(1, 0)][2] & TRUNC
    assert not w.env_flags() & 3
    st = w.state()
    assert (1, 0) not in st and (1, 1) in st and not rows[(1, 1)][2] & (TERM | TRUNC)
    # the predator gained the prey's energy: 5.0 - 0.2 + (3.0 - 0.05)  (ECO:786-883, intake cap inf)
    assert st[PRED]["energy"] == pytest.approx(5.0 - 0.2 + 3.0 - 0.05)
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_time_limit_truncates_with_final_bootstrap_observations(backend):
    """REF:296-329 — max_steps = 1: both agents truncated (not terminated) with an observation of the right shape"""
    w = World(backend, dict(NO_BIRTHS, max_steps=1), [cell(1, 1), cell(G - 2, G - 2)] + FAR_GRASS, [1.0, 1.0])
    rows = w.step({PRED: STAY, PREY: STAY})
    assert set(rows) == {PRED, PREY}
    assert rows[PRED][0].shape == (4, 7, 7) and rows[PREY][0].shape == (4, 9, 9)
    for k in (PRED, PREY):
        assert rows[k][2] & TRUNC and not rows[k][2] & TERM
    assert w.env_flags() & 2 and not w.env_flags() & 1
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_offspring_inherits_mutated_parent_genome_and_fixed_initial_energy(backend):
    """REF:417-446 — threshold 10, founder speed exactly 1.0, mutation rate 1.0 / std 0.01: the child starts with
    initial_energy_predator, the parent pays exactly that, the child's speed is the parent's plus the mutation draw"""
    ov = dict(predator_creation_energy_threshold=10.0, prey_creation_energy_threshold=999.0, initial_energy_predator=20.25,
              energy_loss_per_step_predator=0.25, genome_mutation={"rate": 1.0, "std": 0.01},
              founder_genome={"predator": {"speed_mean": 1.0, "speed_std": 0.0}})
    w = World(backend, ov, [cell(5, 5), cell(20, 20)] + FAR_GRASS, [1.0, 1.0], reals=[0.5, 0.00390625])  # u < rate, then delta
    rows = w.step({PRED: STAY, PREY: STAY})
    st = w.state()
    child = (0, 1)
    assert child in st and child in rows  # newborn: first unused id (ECO:260-272), observed in the same call
    assert st[child]["energy"] == 20.25
    assert st[PRED]["energy"] == (20.25 - 0.25) - 20.25
    assert st[PRED]["speed"] == 1.0 and st[child]["speed"] == 1.0 + 0.00390625
    assert rows[PRED][1] == 10.0  # reproduction_reward_predator
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("energy,births", [(9.25, 0), (10.25, 1)])
def test_reproduction_threshold_uses_fixed_base_threshold(backend, energy, births):
    """REF:449-472 — energy 9.0 at the reproduction phase: no child; 10.0: one child (threshold 10.0, no mutation)"""
    ov = dict(predator_creation_energy_threshold=10.0, prey_creation_energy_threshold=999.0, initial_energy_predator=energy,
              energy_loss_per_step_predator=0.25, genome_mutation={"rate": 0.0, "std": 0.0},
              founder_genome={"predator": {"speed_mean": 1.0, "speed_std": 0.0}})
    w = World(backend, ov, [cell(5, 5), cell(20, 20)] + FAR_GRASS, [1.0, 1.0])
    w.step({PRED: STAY, PREY: STAY})
    assert sum(1 for k in w.state() if k[0] == 0) == 1 + births
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_observation_edges_are_clipped_and_zero_padded(backend):
    """REF:482-495 — predator on (0, 0): the window rows / columns outside the grid are 0 in every grid channel"""
    w = World(backend, {}, [cell(0, 0), cell(20, 20)] + FAR_GRASS, [1.0, 1.0])
    obs = w.rows()[PRED][0]
    assert obs.shape == (4, 7, 7)
    assert np.all(obs[:3, :3, :] == 0.0) and np.all(obs[:3, :, :3] == 0.0)
    assert obs[0, 3, 3] == np.float32(5.0)  # itself, at the window centre, in the predator channel
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("speed,dest", [(1.0, (11, 10)), (1.6, (12, 10))])
def test_speed_gates_the_distance_two_move(backend, speed, dest):
    """REF:498-513 slow (speed 1.0 < 1.5) clips (2, 0) to (1, 0); REF:516-531 fast (1.6) moves two cells"""
    w = World(backend, dict(NO_BIRTHS, speed_distance_threshold=1.5), [cell(10, 10), cell(20, 20)] + FAR_GRASS, [speed, 1.0])
    w.step({PRED: MOVE_2_0, PREY: STAY})
    assert w.state()[PRED]["xy"] == dest
    w.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("action,cost", [(STAY, 0.0), (MOVE_2_0, 0.05 * 2.0 * (2.0 ** 2))])
def test_movement_cost_uses_actual_distance_and_superlinear_speed(backend, action, cost):
    """REF:534-557 stationary at speed 2.0 pays only the basal 0.2; REF:560-586 a two-cell move at speed 2.0 costs
    0.05 * 2 * 2**2 on top"""
    ov = dict(NO_BIRTHS, energy_loss_per_step_predator=0.2, movement_energy_cost_per_cell_predator=0.05, movement_speed_cost_exponent=2.0,
              initial_energy_predator=10.0)
    w = World(backend, ov, [cell(10, 10), cell(20, 20)] + FAR_GRASS, [2.0, 1.0])
    w.step({PRED: action, PREY: STAY})
    st = w.state()[PRED]
    assert st["energy"] == pytest.approx(10.0 - 0.2 - cost)
    assert st["xy"] == ((12, 10) if action == MOVE_2_0 else (10, 10))
    w.close()


def test_action_table_is_the_extended_moore_neighbourhood():
    """REF:475-479 — 25 actions, (2, 0) among them (ECO:225-232)"""
    table = {a: (a // 5 - 2, a % 5 - 2) for a in range(ECO_CONFIG["action_range"] ** 2)}
    assert len(table) == 25 and table[MOVE_2_0] == (2, 0) and table[STAY] == (0, 0)
