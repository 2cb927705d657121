"""ECO (eco_evolutionary env) on the GPU vs the CPU oracle and vs the reference's golden trajectories (`-m gpu`).

Everything goes through the C-ABI.  Bit-exact: ids, row layout, flags, positions, float64 energies, ages, genome speeds,
dead_prey, the drifting active_num_* counters, float32 observations and rewards.  The locomotion cost's
`speed ** exponent` is glibc's pow on both sides: libm in the oracle (as in CPython), include/ppg_pow.h on the device."""
import numpy as np
import pytest

from predpreygrass_b200.config import ECO_CONFIG, VARIANT_ECO, make_config
from tests.helpers import config_from_golden, golden_cases, id_order_rows, load_golden, sha_f32
from tests.parity import lockstep_parity

pytestmark = pytest.mark.gpu

CROWDED = dict(ECO_CONFIG, grid_size=8, initial_num_grass=24, n_initial_active_predators=6, n_initial_active_prey=14,
               predator_creation_energy_threshold=6.0, prey_creation_energy_threshold=3.5, initial_energy_predator=3.0,
               initial_energy_prey=2.0, energy_gain_per_step_grass=0.5, energy_loss_per_step_prey=0.02,
               energy_loss_per_step_predator=0.1, predator_obs_range=5, prey_obs_range=7, n_possible_predators=300,
               n_possible_prey=400, max_steps=60, genome_mutation={"rate": 0.5, "std": 0.4})
RICH = dict(ECO_CONFIG, energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0, predator_creation_energy_threshold=8.0,
            energy_loss_per_step_predator=0.1)


def eco(cfg, **kw):
    return make_config(cfg, variant=VARIANT_ECO, **kw)


def test_eco_default_philox():
    """BASELINE configs[3] world: default ECO config, Philox trait / placement streams, several episodes per env."""
    st = lockstep_parity(eco(ECO_CONFIG, cap_live=(64, 128), seed=7), 256, 160, state_envs=(0, 5, 255))
    assert st["status_envs"] == 0 and st["episodes"] > 256 and st["births_prey"] > 0


def test_eco_reproduction_heavy():
    """births, mutation draws and slot recycling dominate (thresholds lowered, grass regrows fast)"""
    st = lockstep_parity(eco(RICH, cap_live=(128, 320), seed=3), 128, 150, state_envs=(0, 64, 127))
    assert st["births_prey"] > 10000 and st["births_pred"] > 100


def test_eco_crowded_spawn_fallback_and_mutation():
    st = lockstep_parity(eco(CROWDED, cap_live=(64, 64), seed=5), 256, 120, state_envs=(0, 17, 255))
    assert st["spawn_fallback"] > 0 and st["births_prey"] > 0


def test_eco_carcasses_and_intake_caps():
    cfg = dict(CROWDED, max_energy_gain_per_prey=0.8, max_energy_gain_per_grass=1.0)
    st = lockstep_parity(eco(cfg, cap_live=(64, 64), seed=9), 256, 120, state_envs=(0, 100))
    assert st["eaten_prey"] > 0


def test_eco_age_caps_and_juvenile_predators():
    cfg = dict(CROWDED, grid_size=10, max_agent_age={"predator": 25, "prey": 12}, carcass_only_predator_age={"predator": 6})
    st = lockstep_parity(eco(cfg, cap_live=(64, 96), seed=11), 256, 120, state_envs=(0, 31))
    assert st["episodes"] > 0


def test_eco_without_genome_three_channel_rows():
    """genome disabled, no speed plane: rows are (3,R,R) and take the generic row writer"""
    cfg = dict(RICH, genome_enabled=False, include_speed_in_obs=False, max_steps=50)
    st = lockstep_parity(eco(cfg, cap_live=(128, 320), seed=13), 64, 120, state_envs=(0, 63))
    assert st["truncated"] > 0


def test_eco_slot_overflow_id_pool_and_idle():
    cfg = dict(CROWDED, n_possible_prey=60, n_possible_predators=16, max_steps=40)
    st = lockstep_parity(eco(cfg, cap_live=(32, 32), seed=2, autoreset=False), 64, 70)
    assert st["births_prey"] > 0


def test_eco_ghost_cells():
    """Prey that age out and are bitten under a finite intake cap in the same step leave a stale value on the reference's
    grid (ECO:826-832 after :1060-1090, removal ECO:332-351 does not zero it); the device carries it as a ghost cell."""
    cfg = dict(CROWDED, max_agent_age={"predator": None, "prey": 9}, max_energy_gain_per_prey=0.4, max_steps=80,
               energy_gain_per_step_grass=0.8, prey_creation_energy_threshold=3.0)
    st = lockstep_parity(eco(cfg, cap_live=(96, 192), seed=17), 256, 100, state_envs=(0, 255))
    assert st["eaten_prey"] > 0  # (status words are compared env by env inside lockstep_parity: a ghost overflow would differ)


def test_eco_lineage_survival_rewards():
    """non-zero lineage_reward_coeff (ECO:943-984): ancestors are paid for births and charged for deaths of their descendants
    (carcass bites and age-outs included); the per-id lineage tables live in HBM"""
    cfg = dict(CROWDED, grid_size=10, lineage_reward_coeff={"predator": 0.5, "prey": -0.25}, max_agent_age={"predator": 25, "prey": 12},
               max_energy_gain_per_prey=0.8)
    st = lockstep_parity(eco(cfg, cap_live=(96, 192), seed=23), 256, 120, state_envs=(0, 255))
    assert st["births_prey"] > 1000 and st["eaten_prey"] > 0
    st = lockstep_parity(eco(dict(RICH, lineage_reward_coeff=0.75), cap_live=(128, 320), seed=29), 128, 150, state_envs=(0, 127))
    assert st["births_prey"] > 1000


def test_eco_4096_envs():
    st = lockstep_parity(eco(ECO_CONFIG, cap_live=(64, 128), seed=21), 4096, 80, state_envs=(0, 4095), check_every=4)
    assert st["status_envs"] == 0


@pytest.mark.parametrize("name", golden_cases(("eco",)))
def test_eco_golden_trajectories_on_gpu(name):
    """Golden trajectories of the unmodified reference ECO class replayed on the GPU (one env, tape-driven)."""
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    z, cfg = load_golden(name)
    c = config_from_golden(cfg, autoreset=False)
    c.cap_live[0] = min(c.cap_live[0], 224)
    c.cap_live[1] = min(c.cap_live[1], 416)
    g = BatchedPredPreyGrass(c, 1)
    g.load_tape([np.concatenate([z["init_cells"], z["fallback_cells"]])], [np.concatenate([z["founder_speed"], z["step_reals"]])])
    g.reset()
    out = g.outputs_numpy()
    rows = id_order_rows(out)
    assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["reset_row_id"])
    assert np.array_equal(sha_f32([out[f"obs{s}"][r] for s, r in rows]), z["reset_sha"])
    T = len(z["steps"])
    for t in range(T):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        act, rank, seen = {}, {}, [0, 0]
        for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1]):
            act[(int(s), int(i))] = int(v)
            rank[(int(s), int(i))] = seen[int(s)]
            seen[int(s)] += 1
        orders = []
        for s in range(2):
            n = out["n"][s]
            a = np.full(max(n, 1), 12, np.int32)
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
            assert np.array_equal(st["energy"][s], z["st_e"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["age"][s], z["st_age"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["speed"][s], z["st_speed"][s0:s1][m]), (name, t, s)
        assert np.array_equal(st["dead_prey"], z["st_dead"][s0:s1][z["st_s"][s0:s1] == 1]), (name, t)
        assert np.array_equal(st["grass_energy"], z["grass_e"][t]), (name, t)
    assert int(out["env_status"][0]) == 0
    g.close()


def test_eco_step_host_and_snapshot_restore():
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    cfg = eco(RICH, cap_live=(128, 320), seed=17)
    a = BatchedPredPreyGrass(cfg, 64)
    b = BatchedPredPreyGrass(cfg, 64)
    a.reset(); b.reset()
    host = b.make_host_buffers()
    for t in range(40):
        a0, a1 = a.random_actions(99)
        a.step(a0, a1)
        n = a.out.n_rows.tolist()
        b0, b1 = b.random_actions(99)
        torch.cuda.synchronize()
        host["actions0"][: len(b0)].copy_(b0.cpu()); host["actions1"][: len(b1)].copy_(b1.cpu())
        n0, n1 = b.step_host(host)
        assert (n0, n1) == (n[0] + n[2], n[1] + n[3])
        for s, k in ((0, n0), (1, n1)):
            assert torch.equal(host[f"obs{s}"][:k], a.out.obs[s][:k].cpu())
            assert torch.equal(host[f"flags{s}"][:k], a.out.flags[s][:k].cpu())
    blob = a.snapshot()
    ref = []
    for t in range(10):
        a0, a1 = a.random_actions(5)
        a.step(a0, a1)
        ref.append(a.outputs_numpy())
    a.restore(blob)
    for t in range(10):
        a0, a1 = a.random_actions(5)
        a.step(a0, a1)
        o = a.outputs_numpy()
        for s in range(2):
            assert np.array_equal(o[f"row_agent{s}"], ref[t][f"row_agent{s}"]), t
            assert np.array_equal(o[f"obs{s}"], ref[t][f"obs{s}"]), t
            assert np.array_equal(o[f"reward{s}"], ref[t][f"reward{s}"]), t
    a.close(); b.close()
