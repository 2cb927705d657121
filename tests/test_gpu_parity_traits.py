"""The other heritable-trait variants of eco_evolutionary on the GPU (`-m gpu`): metabolic_rate (MR), offspring
investment fraction (INV), cooperation rate (COOP) — the trait_mode branches of csrc/ppg_eco.cu vs the CPU oracle (Philox
streams, hundreds of envs) and vs golden trajectories of the unmodified reference classes (one env, tape-driven).

Bit-exact: founder counts, ids, row layout, flags, positions, float64 energies, ages, trait values, float32 observations
and rewards, the active_num_* counters.  `metabolic_rate ** alpha` is glibc's pow on both sides."""
import numpy as np
import pytest

from predpreygrass_b200.config import (CADENCE_CONFIG, COOPERATION_CONFIG, INVESTMENT_CONFIG, METABOLIC_CONFIG, VARIANT_ECO,
                                      make_config)
from tests.helpers import config_from_golden, golden_cases, id_order_rows, load_golden, sha_f32
from tests.parity import lockstep_parity

pytestmark = pytest.mark.gpu

CROWDED = dict(grid_size=9, initial_num_grass=24, n_initial_active_predators=6, n_initial_active_prey=14,
               predator_creation_energy_threshold=6.0, prey_creation_energy_threshold=4.5, energy_gain_per_step_grass=0.3,
               predator_obs_range=5, prey_obs_range=7, n_possible_predators=300, n_possible_prey=400, max_steps=80)
RICH = dict(energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0, predator_creation_energy_threshold=8.0)


def trait(cfg, **kw):
    return make_config(cfg, variant=VARIANT_ECO, **kw)


def test_mr_default_philox():
    """default metabolic_rate world: random founder counts per episode, rate-scaled decay and gains, satiation cooldown 8"""
    st = lockstep_parity(trait(METABOLIC_CONFIG, cap_live=(64, 160), seed=7), 256, 200, state_envs=(0, 5, 255))
    assert st["status_envs"] == 0 and st["episodes"] > 256 and st["eaten_prey"] > 0


def test_mr_rich_mutation_density_cap():
    cfg = dict(METABOLIC_CONFIG, **RICH, basal_energy_cost_predator=0.1, genome_mutation={"rate": 1.0, "std": 0.3},
               metabolic_rate_alpha=0.7, predator_satiation_cooldown=3, predator_reproduction_max_ratio=0.4, max_steps=150)
    st = lockstep_parity(trait(cfg, cap_live=(128, 320), seed=3), 128, 170, state_envs=(0, 64, 127))
    assert st["births_prey"] > 500 and st["births_pred"] > 0


def test_mr_crowded_spawn_fallback():
    cfg = dict(METABOLIC_CONFIG, **CROWDED, basal_energy_cost_prey=0.02, basal_energy_cost_predator=0.1, initial_energy_predator=3.0,
               initial_energy_prey=2.0, max_energy_gain_per_prey=1.5, movement_energy_cost_per_cell_prey=0.03,
               movement_energy_cost_per_cell_predator=0.05, genome_mutation={"rate": 0.5, "std": 0.2},
               n_initial_active_predators_min=2, n_initial_active_prey_min=5)
    st = lockstep_parity(trait(cfg, cap_live=(64, 81), seed=5), 256, 120, state_envs=(0, 17, 255))
    assert st["births_prey"] > 0 and st["episodes"] > 256


def test_inv_default_philox():
    st = lockstep_parity(trait(INVESTMENT_CONFIG, cap_live=(64, 160), seed=11), 256, 200, state_envs=(0, 9, 255))
    assert st["status_envs"] == 0 and st["episodes"] > 256


def test_inv_rich_mutation():
    cfg = dict(INVESTMENT_CONFIG, **RICH, energy_loss_per_step_predator=0.1, genome_mutation={"rate": 0.6, "std": 0.15},
               predator_satiation_cooldown=2, max_steps=150)
    st = lockstep_parity(trait(cfg, cap_live=(128, 384), seed=13), 128, 170, state_envs=(0, 127))
    assert st["births_prey"] > 500


def test_inv_without_genome():
    cfg = dict(INVESTMENT_CONFIG, **RICH, genome_enabled=False, energy_loss_per_step_predator=0.1, max_steps=60)
    st = lockstep_parity(trait(cfg, cap_live=(128, 384), seed=2), 64, 130, state_envs=(0, 63))
    assert st["truncated"] > 0


def test_coop_default_philox():
    st = lockstep_parity(trait(COOPERATION_CONFIG, cap_live=(64, 160), seed=17), 256, 200, state_envs=(0, 3, 255))
    assert st["status_envs"] == 0 and st["episodes"] > 256


def test_coop_sharing_dominates():
    """large cooperation rates and radius: almost every meal is split among neighbours (COOP:537-589)"""
    fg = {"predator": {"cooperation_rate_mean": 0.5, "cooperation_rate_std": 0.3},
          "prey": {"cooperation_rate_mean": 0.5, "cooperation_rate_std": 0.3}}
    cfg = dict(COOPERATION_CONFIG, **CROWDED, basal_energy_cost_prey=0.02, basal_energy_cost_predator=0.1,
               initial_energy_predator=3.0, initial_energy_prey=2.0, cooperation_range=2,
               genome_mutation={"rate": 0.5, "std": 0.3}, founder_genome=fg)
    st = lockstep_parity(trait(cfg, cap_live=(64, 81), seed=19), 256, 120, state_envs=(0, 100, 255))
    assert st["eaten_prey"] > 0 and st["grass_eaten"] > 1000


def test_cad_default_philox():
    """default cadence world: move accumulators, speed-scaled basal cost, radius-1 catches, random spawn neighbour, action mask bit"""
    st = lockstep_parity(trait(CADENCE_CONFIG, cap_live=(64, 160), seed=29), 256, 200, state_envs=(0, 7, 255))
    assert st["status_envs"] == 0 and st["episodes"] > 256 and st["eaten_prey"] > 0


def test_cad_rich_mutation_and_age_caps():
    cfg = dict(CADENCE_CONFIG, **RICH, energy_loss_per_step_predator=0.1, genome_mutation={"rate": 0.6, "std": 0.15},
               max_agent_age={"predator": 40, "prey": 25}, max_cooldown=4, max_energy_gain_per_grass=1.0, max_steps=150)
    st = lockstep_parity(trait(cfg, cap_live=(128, 384), seed=31), 128, 170, state_envs=(0, 64, 127))
    assert st["births_prey"] > 500 and st["births_pred"] > 0


def test_cad_crowded_spawn_choices():
    cfg = dict(CADENCE_CONFIG, **CROWDED, energy_loss_per_step_prey=0.02, energy_loss_per_step_predator=0.1, initial_energy_predator=3.0,
               initial_energy_prey=2.0, max_cooldown=3, genome_mutation={"rate": 0.5, "std": 0.3})
    st = lockstep_parity(trait(cfg, cap_live=(64, 81), seed=37), 256, 120, state_envs=(0, 17, 255))
    assert st["births_prey"] > 0 and st["episodes"] > 256


def test_cad_without_genome():
    cfg = dict(CADENCE_CONFIG, **RICH, genome_enabled=False, include_speed_in_obs=False, energy_loss_per_step_predator=0.1, max_steps=60)
    st = lockstep_parity(trait(cfg, cap_live=(128, 384), seed=41), 64, 130, state_envs=(0, 63))
    assert st["episodes"] > 0


@pytest.mark.parametrize("base", ["mr", "inv", "coop"])
def test_trait_episode_event_counters(base):
    """the counters behind `training_metrics` (births blocked by an exhausted id pool / the density cap, catches blocked by
    satiation, COOP's donated energy; MR:1347-1350, COOP:1365-1368) against the oracle, every step, on worlds that hit them:
    tiny id pools, a tight density cap, a long satiation cooldown, generous sharing"""
    cfg = dict({"mr": METABOLIC_CONFIG, "inv": INVESTMENT_CONFIG, "coop": COOPERATION_CONFIG}[base], **RICH,
               n_possible_predators=24, n_possible_prey=60, genome_mutation={"rate": 1.0, "std": 0.3}, max_steps=120)
    if base == "mr":
        cfg.update(predator_reproduction_max_ratio=0.3, predator_satiation_cooldown=6)
    if base == "inv":
        cfg.update(predator_satiation_cooldown=6)
    if base == "coop":
        cfg.update(cooperation_range=3)
    envs = tuple(range(0, 64, 7))
    st = lockstep_parity(trait(cfg, cap_live=(32, 64), seed=43, track_episode_sums=True), 64, 140, state_envs=envs)
    assert st["births_prey"] > 0 and st["status_or"] & 0x10  # PPG_STATUS_ID_POOL_EMPTY: some births were blocked by capacity


def test_traits_2048_envs():
    for base, seed in ((METABOLIC_CONFIG, 21), (INVESTMENT_CONFIG, 22), (COOPERATION_CONFIG, 23), (CADENCE_CONFIG, 24)):
        st = lockstep_parity(trait(base, cap_live=(64, 160), seed=seed), 2048, 80, state_envs=(0, 2047), check_every=4)
        assert st["status_envs"] == 0


@pytest.mark.parametrize("name", golden_cases(("mr", "inv", "coop", "cad")))
def test_trait_golden_trajectories_on_gpu(name):
    """Golden trajectories of the unmodified reference classes replayed on the GPU (one env, tape-driven: founder counts,
    cells, trait values, mutation draws, spawn-fallback cells all from the recording)."""
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    z, cfg = load_golden(name)
    c = config_from_golden(cfg, autoreset=False)
    g = BatchedPredPreyGrass(c, 1)
    stay = (cfg["action_range"] ** 2) // 2
    cad = cfg["variant"] == "cad"
    if cad:  # fixed number of founders; per founder a speed and an accumulator phase; one recorded cell per birth
        g.load_tape([np.concatenate([z["init_cells"], z["fallback_cells"]])],
                    [np.concatenate([z["founder_trait"], z["founder_acc"], z["step_reals"]])])
    else:
        g.load_tape([np.concatenate([z["n_found"], z["init_cells"], z["fallback_cells"]])],
                    [np.concatenate([z["founder_trait"], z["step_reals"]])])
    g.reset()
    out = g.outputs_numpy()
    rows = id_order_rows(out)
    assert [s for s, _ in rows] == list(z["reset_row_s"])
    if cad:
        assert [int(out[f"flags{s}"][r]) >> 7 for s, r in rows] == list(z["reset_frozen"])
    assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["reset_row_id"])
    assert np.array_equal(sha_f32([out[f"obs{s}"][r] for s, r in rows]), z["reset_sha"])
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        act, rank, seen = {}, {}, [0, 0]
        for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1]):
            act[(int(s), int(i))] = int(v)
            rank[(int(s), int(i))] = seen[int(s)]
            seen[int(s)] += 1
        orders = []
        for s in range(2):
            n = out["n"][s]
            a = np.full(max(n, 1), stay, np.int32)
            o = np.zeros(max(n, 1), np.int32)
            for r in range(n):
                if not (out[f"flags{s}"][r] & 1):
                    a[r] = act[(s, int(out[f"row_agent{s}"][r]))]
                    o[r] = rank[(s, int(out[f"row_agent{s}"][r]))]
            g.actions[s][: len(a)].copy_(torch.from_numpy(a))
            orders.append(torch.from_numpy(o).cuda())
        if str(z["order"]) == "shuffle":
            g.step_ordered(g.actions[0], g.actions[1], orders[0], orders[1])
        else:
            g.step()
        out = g.outputs_numpy()
        rows = id_order_rows(out)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        assert [s for s, _ in rows] == list(z["row_s"][r0:r1]), (name, t)
        assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["row_id"][r0:r1]), (name, t)
        rew = np.array([out[f"reward{s}"][r] for s, r in rows], np.float32)
        assert np.array_equal(rew, z["row_rew"][r0:r1].astype(np.float32)), (name, t)
        fl = np.array([out[f"flags{s}"][r] for s, r in rows], np.uint8)
        assert np.array_equal(fl & 1, z["row_term"][r0:r1]), (name, t)
        assert np.array_equal((fl >> 1) & 1, z["row_trunc"][r0:r1]), (name, t)
        if cad:
            assert np.array_equal(fl >> 7, z["row_frozen"][r0:r1]), (name, t)
        assert np.array_equal(sha_f32([out[f"obs{s}"][r] for s, r in rows]), z["obs_sha"][t]), (name, t)
        assert bool(out["env_flags"][0] & 1) == bool(z["all_term"][t]), (name, t)
        assert bool(out["env_flags"][0] & 2) == bool(z["all_trunc"][t]), (name, t)
        assert list(out["env_count"][0]) == list(z["active"][t]), (name, t)
        if out["env_flags"][0] & 3:
            break
        st = g.read_env_eco(0)
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        for s in range(2):
            m = z["st_s"][s0:s1] == s
            assert np.array_equal(st["ids"][s], z["st_id"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["xy"][s][:, 0], z["st_x"][s0:s1][m]) and np.array_equal(st["xy"][s][:, 1], z["st_y"][s0:s1][m])
            assert np.array_equal(st["energy"][s], z["st_e"][s0:s1][m]), (name, t, s, st["energy"][s] - z["st_e"][s0:s1][m])
            assert np.array_equal(st["age"][s], z["st_age"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["speed"][s], z["st_trait"][s0:s1][m]), (name, t, s)
            if cad:
                assert np.array_equal(g.read_env_acc(0)[s], z["st_acc"][s0:s1][m]), (name, t, s)
        assert np.array_equal(st["grass_energy"], z["grass_e"][t]), (name, t)
    assert int(out["env_status"][0]) & ~0x04 == 0  # only PPG_STATUS_TAPE_EXHAUSTED may be set (recordings cut before the end)
    g.close()


def test_trait_snapshot_restore():
    from predpreygrass_b200.batched import BatchedPredPreyGrass

    a = BatchedPredPreyGrass(trait(dict(METABOLIC_CONFIG, **RICH), cap_live=(128, 320), seed=17), 64)
    a.reset()
    for t in range(60):
        a0, a1 = a.random_actions(99)
        a.step(a0, a1)
    blob = a.snapshot()
    ref = []
    for t in range(12):
        a0, a1 = a.random_actions(5)
        a.step(a0, a1)
        ref.append(a.outputs_numpy())
    a.restore(blob)
    for t in range(12):
        a0, a1 = a.random_actions(5)
        a.step(a0, a1)
        o = a.outputs_numpy()
        for s in range(2):
            assert np.array_equal(o[f"row_agent{s}"], ref[t][f"row_agent{s}"]), t
            assert np.array_equal(o[f"obs{s}"], ref[t][f"obs{s}"]), t
            assert np.array_equal(o[f"reward{s}"], ref[t][f"reward{s}"]), t
    a.close()
