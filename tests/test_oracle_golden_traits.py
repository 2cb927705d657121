"""The oracle's trait-variant branches (metabolic_rate, offspring_investment_fraction, cooperation_rate) against
trajectories recorded from the unmodified reference classes (tests/golden/{mr,inv,coop}_*.npz, make_golden_traits.py).

Bit-exact on everything: founder counts, ids, positions, float64 energies, ages, trait values, rewards, flags, the float32
observation bytes (sha1 + full arrays of sampled steps), the float32 grid (sha1), the active_num_* counters."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.helpers import config_from_golden, golden_cases, id_order_rows, load_golden, sha_f32


@pytest.mark.parametrize("name", golden_cases(("mr", "inv", "coop", "cad")))
def test_trait_oracle_replays_reference(name):
    z, cfg = load_golden(name)
    c = config_from_golden(cfg, autoreset=False)
    o = Oracle(c, 1)
    o.load_tape([z["fallback_cells"]], [z["step_reals"]])
    cad = cfg["variant"] == "cad"
    out = o.env_reset_trait(0, int(z["n_found"][0]), int(z["n_found"][1]), z["init_cells"],
                            np.concatenate([z["founder_trait"], z["founder_acc"]]) if cad else z["founder_trait"])
    rows = id_order_rows(out)
    if cad:  # the action mask of the reset observations (CAD:746-753) as the PPG_ROW_FROZEN bit
        assert [int(out[f"flags{s}"][r]) >> 7 for s, r in rows] == list(z["reset_frozen"])
    assert [s for s, _ in rows] == list(z["reset_row_s"])
    assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["reset_row_id"])
    assert np.array_equal(sha_f32([out[f"obs{s}"][r] for s, r in rows]), z["reset_sha"])

    full = set(int(t) for t in z["full_obs_steps"])
    for t in range(len(z["steps"])):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        out = o.env_step_ordered(0, z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])
        rows = id_order_rows(out)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        assert [s for s, _ in rows] == list(z["row_s"][r0:r1]), (name, t)
        assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["row_id"][r0:r1]), (name, t)
        rew = np.array([out[f"reward64_{s}"][r] for s, r in rows])
        assert np.array_equal(rew, z["row_rew"][r0:r1]), (name, t)
        fl = np.array([out[f"flags{s}"][r] for s, r in rows], np.uint8)
        assert np.array_equal(fl & 1, z["row_term"][r0:r1]), (name, t)
        assert np.array_equal((fl >> 1) & 1, z["row_trunc"][r0:r1]), (name, t)
        if cad:
            assert np.array_equal(fl >> 7, z["row_frozen"][r0:r1]), (name, t)
        if t in full:
            for s in range(2):
                mine = [out[f"obs{s}"][r] for ss, r in rows if ss == s]
                ref = z[f"full_obs_{t}_{s}"]
                assert len(mine) == len(ref), (name, t, s)
                for k in range(len(mine)):
                    assert np.array_equal(mine[k], ref[k]), (name, t, s, k, np.argwhere(mine[k] != ref[k])[:4])
        assert np.array_equal(sha_f32([out[f"obs{s}"][r] for s, r in rows]), z["obs_sha"][t]), (name, t)
        assert bool(out["env_flags"][0] & 1) == bool(z["all_term"][t]), (name, t)
        assert bool(out["env_flags"][0] & 2) == bool(z["all_trunc"][t]), (name, t)
        assert int(out["env_step"][0]) == int(z["steps"][t])
        assert list(out["env_count"][0]) == list(z["active"][t]), (name, t)
        if z["all_term"][t] or z["all_trunc"][t]:
            break  # the reference clears self.agents; the remaining state is dead
        st = o.read_env_eco(0)
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        for s in range(2):
            m = z["st_s"][s0:s1] == s
            assert np.array_equal(st["ids"][s], z["st_id"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["xy"][s][:, 0], z["st_x"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["xy"][s][:, 1], z["st_y"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["energy"][s], z["st_e"][s0:s1][m]), (name, t, s, st["energy"][s] - z["st_e"][s0:s1][m])
            assert np.array_equal(st["age"][s], z["st_age"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["speed"][s], z["st_trait"][s0:s1][m]), (name, t, s)
            if cad:
                assert np.array_equal(o.read_env_acc(0)[s], z["st_acc"][s0:s1][m]), (name, t, s)
        assert np.array_equal(st["grass_energy"], z["grass_e"][t]), (name, t)
        assert np.array_equal(sha_f32([o.read_grid(0)]), z["grid_sha"][t]), (name, t)
        g0, g1 = z["ag_off"][t], z["ag_off"][t + 1]
        ags, agi = o.env_agents(0)
        assert list(ags) == list(z["ag_s"][g0:g1]) and list(agi) == list(z["ag_id"][g0:g1]), (name, t)
    assert int(out["env_status"][0]) & ~0x04 == 0  # only PPG_STATUS_TAPE_EXHAUSTED may be set (recordings cut before the end)
    o.close()
