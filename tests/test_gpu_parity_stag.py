"""STAG (stag_hunt_forward_view_nature_nurture env) on the GPU vs the CPU oracle and vs the reference's golden
trajectories (`-m gpu`).

Everything goes through the C-ABI.  Bit-exact: ids, row layout, flags, positions, float64 energies, ages, facings,
cooperation traits, team-capture counters, float32 observations (forward-shifted predator windows included) and
rewards.  The capture success probability uses include/ppg_philox.h's `ppg_pow_frac` on both sides of the lockstep
tests; against the golden files (CPython's libm pow) `last_success_prob` is compared to 1e-12 relative."""
import numpy as np
import pytest

from predpreygrass_b200.config import STAG_CONFIG, VARIANT_STAG, make_config
from tests.helpers import config_from_golden, dict_order_rows, golden_cases, load_golden, sha_f32
from tests.parity import lockstep_parity

pytestmark = pytest.mark.gpu

# a small, crowded world: captures, births and blocked moves every step; both predator types; proportional split
CROWDED = dict(STAG_CONFIG, grid_size=9, initial_num_grass=30, predator_obs_range=7, prey_obs_range=5, max_steps=70,
               n_possible_type_1_predators=150, n_possible_type_2_predators=60, n_possible_type_1_prey=120, n_possible_type_2_prey=200,
               n_initial_active_type_1_predator=5, n_initial_active_type_2_predator=2, n_initial_active_type_1_prey=5,
               n_initial_active_type_2_prey=9, energy_treshold_creation_predator=6.0,
               energy_treshold_creation_prey={"type_1_prey": 11.5, "type_2_prey": 2.1}, energy_gain_per_step_grass=0.4,
               reproduction_reward_predator={"type_1_predator": 10.0, "type_2_predator": 3.0}, team_capture_equal_split=False,
               team_capture_join_cost=0.3, death_penalty_predator=-1.0, death_penalty_type_1_prey=-2.0, death_penalty_type_2_prey=-0.5)
RICH = dict(STAG_CONFIG, energy_gain_per_step_grass=0.5, energy_treshold_creation_predator=7.0,
            energy_treshold_creation_prey={"type_1_prey": 13.0, "type_2_prey": 2.2})


def stag(cfg, **kw):
    return make_config(cfg, variant=VARIANT_STAG, **kw)


def test_stag_default_philox():
    """BASELINE configs[4] world: default STAG config, Philox placement / facing / trait / capture streams."""
    st = lockstep_parity(stag(STAG_CONFIG, cap_live=(64, 160), seed=7), 256, 200, state_envs=(0, 5, 255))
    print(st)
    assert st["status_or"] in (0, 1) and st["capture_attempts"] > 1000 and st["eaten_prey"] > 0 and st["births_prey"] > 0


def test_stag_reproduction_heavy():
    st = lockstep_parity(stag(RICH, cap_live=(128, 384), seed=3), 128, 150, state_envs=(0, 64, 127))
    assert st["births_prey"] > 5000 and st["births_pred"] > 100


def test_stag_crowded_two_predator_types_join_cost_and_penalties():
    st = lockstep_parity(stag(CROWDED, cap_live=(96, 96), seed=5), 256, 150, state_envs=(0, 17, 255))
    assert st["capture_attempts"] > 0 and st["births_pred"] > 0 and st["starved_pred"] > 0 and st["episodes"] > 0


def test_stag_crowded_not_strict_deterministic_no_trait():
    cfg = dict(CROWDED, strict_rllib_output=False, team_capture_success_model="deterministic", coop_trait_enabled=False,
               team_capture_scavenger_fraction=0.5)
    st = lockstep_parity(stag(cfg, cap_live=(96, 96), seed=9), 128, 120, state_envs=(0, 100))
    assert st["capture_attempts"] > 0


def test_stag_probabilistic_margin_spawn_fallback():
    cfg = dict(CROWDED, grid_size=6, initial_num_grass=12, team_capture_success_model="probabilistic", team_capture_margin=1.5,
               team_capture_min_success_prob=0.1, team_capture_join_cost=0.0, predator_obs_range=9, prey_obs_range=9)
    st = lockstep_parity(stag(cfg, cap_live=(64, 64), seed=11), 256, 120, state_envs=(0, 31))
    assert st["spawn_fallback"] > 0


def test_stag_slot_overflow_id_pool_and_idle():
    cfg = dict(CROWDED, n_possible_type_2_prey=24, n_possible_type_1_predators=12, max_steps=40)
    st = lockstep_parity(stag(cfg, cap_live=(32, 32), seed=2, autoreset=False), 64, 70)
    assert st["births_prey"] > 0


def test_stag_walls_and_line_of_sight():
    """manual walls (channel 0, blocked moves, spawn and placement exclusion, STAG:860-864,1026-1037,2107-2160) and the
    line-of-sight test of prey moves with a 5x5 action range (STAG:875-925), Philox placement around the walls"""
    walls = [(3, y) for y in range(2, 8)] + [(6, 1), (6, 2), (7, 7), (8, 7), (1, 8), (9, 9)]
    cfg = dict(CROWDED, grid_size=10, manual_wall_positions=walls, respect_los_for_movement=True, type_2_action_range=5)
    st = lockstep_parity(stag(cfg, cap_live=(96, 224), seed=31), 256, 120, state_envs=(0, 100, 255))
    assert st["births_prey"] > 0 and st["episodes"] > 0
    cfg = dict(STAG_CONFIG, manual_wall_positions=[(x, 15) for x in range(5, 25)] + [(10, y) for y in range(3, 12)], respect_los_for_movement=True)
    st = lockstep_parity(stag(cfg, cap_live=(64, 192), seed=37), 128, 100, state_envs=(0, 127))
    assert st["status_envs"] == 0


def test_stag_4096_envs():
    st = lockstep_parity(stag(STAG_CONFIG, cap_live=(64, 160), seed=21), 4096, 80, state_envs=(0, 4095), check_every=4)
    print(st)
    assert st["status_or"] in (0, 1)


def ref_order_rows(out, live_keys):
    rows = dict_order_rows(out)
    by_key = {(s, int(out[f"row_agent{s}"][r])): (s, r) for s, r in rows}
    live = [by_key[k] for k in live_keys]
    ended = sorted((k for k in by_key if k not in set(live_keys)))
    return live + [by_key[k] for k in ended]


@pytest.mark.parametrize("name", golden_cases(("stag",)))
def test_stag_golden_trajectories_on_gpu(name):
    """Golden trajectories of the unmodified reference STAG class replayed on the GPU (one env, tape-driven)."""
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    z, cfg = load_golden(name)
    # the visibility channel is a constant plane of ones in the reference (STAG:408-412,993-994) and not part of the device
    # rows (include/ppg.h): the device runs without it and the plane is appended here, as PredPreyGrassStag does
    vis = bool(cfg.get("include_visibility_channel"))
    c = config_from_golden(dict(cfg, include_visibility_channel=False), autoreset=False)
    c.cap_live[0] = min(c.cap_live[0], 224)
    c.cap_live[1] = min(c.cap_live[1], 416)
    g = BatchedPredPreyGrass(c, 1)

    def with_vis(o, flags=0):
        return np.concatenate([o, np.full((1,) + o.shape[1:], 0.0 if flags & 1 else 1.0, np.float32)]) if vis else o

    reals = np.concatenate([z["founder_trait_raw"] if c.coop_trait_enabled else np.zeros(0), z["step_reals"]])
    g.load_tape([np.concatenate([z["init_cells"], z["founder_facing"], z["step_ints"]])], [reals])
    g.reset()
    out = g.outputs_numpy()
    keys = list(zip(z["reset_row_s"].tolist(), z["reset_row_id"].tolist()))
    rows = ref_order_rows(out, keys)
    assert np.array_equal(sha_f32([with_vis(out[f"obs{s}"][r]) for s, r in rows]), z["reset_sha"])
    T = len(z["steps"])
    full = set(int(t) for t in z["full_obs_steps"])
    for t in range(T):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        # the harness built the action dict from `self.agents`, which still lists the agents that ended in the previous
        # step (strict_rllib_output, STAG:565-573); the reference skips their keys (STAG:806-807), the rank counts live ones
        alive = {(s, int(out[f"row_agent{s}"][r])) for s in range(2) for r in range(out["n"][s]) if not (out[f"flags{s}"][r] & 1)}
        act, rank, seen = {}, {}, [0, 0]
        for s, i, mv, jn in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_move"][a0:a1], z["act_join"][a0:a1]):
            if (int(s), int(i)) not in alive:
                continue
            act[(int(s), int(i))] = int(mv) | (max(int(jn), 0) << 8)
            rank[(int(s), int(i))] = seen[int(s)]
            seen[int(s)] += 1
        orders = []
        for s in range(2):
            n = out["n"][s]
            a = np.full(max(n, 1), 4, np.int32)
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
        g0, g1 = z["ag_off"][t], z["ag_off"][t + 1]
        live_keys = list(zip(z["ag_s"][g0:g1].tolist(), z["ag_id"][g0:g1].tolist()))
        rows = ref_order_rows(out, live_keys)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        assert [s for s, _ in rows] == list(z["row_s"][r0:r1]), (name, t)
        assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["row_id"][r0:r1]), (name, t)
        rew = np.array([out[f"reward{s}"][r] for s, r in rows], np.float32)
        assert np.array_equal(rew, z["row_rew"][r0:r1].astype(np.float32)), (name, t)
        fl = np.array([out[f"flags{s}"][r] for s, r in rows], np.uint8)
        assert np.array_equal(fl & 1, z["row_term"][r0:r1]), (name, t)
        assert np.array_equal((fl >> 1) & 1, z["row_trunc"][r0:r1]), (name, t)
        if t in full:
            for s in range(2):
                mine = [with_vis(out[f"obs{s}"][r], out[f"flags{s}"][r]) for ss, r in rows if ss == s]
                ref = z[f"full_obs_{t}_{s}"]
                assert len(mine) == len(ref), (name, t, s)
                for k in range(len(mine)):
                    assert np.array_equal(mine[k], ref[k]), (name, t, s, k, np.argwhere(mine[k] != ref[k])[:4])
        ordered = ([with_vis(out[f"obs{s}"][r], out[f"flags{s}"][r]) for s, r in rows if s == 0] +
                   [with_vis(out[f"obs{s}"][r], out[f"flags{s}"][r]) for s, r in rows if s == 1])
        assert np.array_equal(sha_f32(ordered), z["obs_sha"][t]), (name, t)
        assert bool(out["env_flags"][0] & 1) == bool(z["all_term"][t]), (name, t)
        assert bool(out["env_flags"][0] & 2) == bool(z["all_trunc"][t]), (name, t)
        assert list(out["env_count"][0]) == list(z["active"][t]), (name, t)
        if out["env_flags"][0] & 3:
            break
        st = g.read_env_stag(0)
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        for s in range(2):
            m = z["st_s"][s0:s1] == s
            assert np.array_equal(st["ids"][s], z["st_id"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["xy"][s][:, 0], z["st_x"][s0:s1][m]) and np.array_equal(st["xy"][s][:, 1], z["st_y"][s0:s1][m])
            assert np.array_equal(st["energy"][s], z["st_e"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["age"][s], z["st_age"][s0:s1][m]), (name, t, s)
        mp = z["st_s"][s0:s1] == 0
        assert np.array_equal(st["facing"], z["st_face"][s0:s1][mp]), (name, t)
        assert np.array_equal(st["trait"], z["st_trait"][s0:s1][mp]), (name, t)
        assert np.array_equal(st["grass_energy"], z["grass_e"][t]), (name, t)
        assert np.array_equal(st["capture"], z["counters"][t]), (name, t, st["capture"], z["counters"][t])
        assert np.allclose(st["capture_real"], z["lastp"][t], rtol=1e-12, atol=0.0), (name, t, st["capture_real"], z["lastp"][t])
    assert int(out["env_status"][0]) == 0
    g.close()


def test_stag_step_host_and_snapshot_restore():
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    cfg = stag(RICH, cap_live=(128, 384), seed=17)
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
