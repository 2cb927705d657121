"""The STAG oracle against trajectories recorded from the unmodified reference (tests/golden/stag_*.npz).

Bit-exact on everything: ids and the order of `self.agents`, positions, float64 energies, ages, facings, cooperation
traits, rewards, flags, the float32 observation bytes (sha1 + full arrays of sampled steps), the float32 grid (sha1),
the team-capture counters and the last success probability / effort ratio (libm pow, like CPython; the device's
bit-exact twin of glibc's pow is pinned in tests/test_pow_port.py)."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.helpers import config_from_golden, dict_order_rows, golden_cases, load_golden, sha_f32


def ref_order_rows(out, live_keys):
    """Rows of env 0 in the order the golden files use: live agents in `self.agents` order, then the ended agents sorted
    by (species, id)."""
    rows = dict_order_rows(out)
    by_key = {(s, int(out[f"row_agent{s}"][r])): (s, r) for s, r in rows}
    live = [by_key[k] for k in live_keys]
    ended = sorted((k for k in by_key if k not in set(live_keys)))
    return live + [by_key[k] for k in ended]


@pytest.mark.parametrize("name", golden_cases(("stag",)))
def test_stag_oracle_replays_reference(name):
    z, cfg = load_golden(name)
    c = config_from_golden(cfg, autoreset=False)
    o = Oracle(c, 1)
    o.load_tape([z["step_ints"]], [z["step_reals"]])
    out = o.env_reset_stag(0, z["init_cells"], z["founder_facing"], z["founder_trait_raw"])
    keys = list(zip(z["reset_row_s"].tolist(), z["reset_row_id"].tolist()))
    rows = ref_order_rows(out, keys)
    assert len(rows) == len(keys)
    assert np.array_equal(sha_f32([out[f"obs{s}"][r] for s, r in rows]), z["reset_sha"])

    T = len(z["steps"])
    full = set(int(t) for t in z["full_obs_steps"])
    for t in range(T):
        a0, a1 = z["act_off"][t], z["act_off"][t + 1]
        av = z["act_move"][a0:a1].astype(np.int32) | (np.maximum(z["act_join"][a0:a1], 0).astype(np.int32) << 8)
        out = o.env_step_ordered(0, z["act_s"][a0:a1], z["act_id"][a0:a1], av)
        g0, g1 = z["ag_off"][t], z["ag_off"][t + 1]
        live_keys = list(zip(z["ag_s"][g0:g1].tolist(), z["ag_id"][g0:g1].tolist()))
        ags, agi = o.env_agents(0)
        assert list(zip(ags.tolist(), agi.tolist())) == live_keys, (name, t)
        rows = ref_order_rows(out, live_keys)
        r0, r1 = z["row_off"][t], z["row_off"][t + 1]
        assert [s for s, _ in rows] == list(z["row_s"][r0:r1]), (name, t)
        assert [int(out[f"row_agent{s}"][r]) for s, r in rows] == list(z["row_id"][r0:r1]), (name, t)
        # the oracle's own row order: per species, agents alive at step start in list order, then newborns in birth order
        for s in range(2):
            mine = [int(out[f"row_agent{ss}"][r]) for ss, r in dict_order_rows(out) if ss == s and not out[f"flags{ss}"][r] & 1]
            assert mine == [i for ss, i in live_keys if ss == s], (name, t, s)
        rew = np.array([out[f"reward64_{s}"][r] for s, r in rows])
        assert np.array_equal(rew, z["row_rew"][r0:r1]), (name, t)
        fl = np.array([out[f"flags{s}"][r] for s, r in rows], np.uint8)
        assert np.array_equal(fl & 1, z["row_term"][r0:r1]), (name, t)
        assert np.array_equal((fl >> 1) & 1, z["row_trunc"][r0:r1]), (name, t)
        if t in full:
            for s in range(2):
                mine = [out[f"obs{s}"][r] for ss, r in rows if ss == s]
                ref = z[f"full_obs_{t}_{s}"]
                assert len(mine) == len(ref), (name, t, s)
                for k in range(len(mine)):
                    assert np.array_equal(mine[k], ref[k]), (name, t, s, k, np.argwhere(mine[k] != ref[k])[:4])
        # the golden sha runs over predators' rows, then prey rows
        ordered = [out[f"obs{s}"][r] for s, r in rows if s == 0] + [out[f"obs{s}"][r] for s, r in rows if s == 1]
        assert np.array_equal(sha_f32(ordered), z["obs_sha"][t]), (name, t)
        assert bool(out["env_flags"][0] & 1) == bool(z["all_term"][t]), (name, t)
        assert bool(out["env_flags"][0] & 2) == bool(z["all_trunc"][t]), (name, t)
        assert int(out["env_step"][0]) == int(z["steps"][t])
        assert list(out["env_count"][0]) == list(z["active"][t]), (name, t)
        st = o.read_env_stag(0)
        s0, s1 = z["st_off"][t], z["st_off"][t + 1]
        for s in range(2):
            m = z["st_s"][s0:s1] == s
            assert np.array_equal(st["ids"][s], z["st_id"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["xy"][s][:, 0], z["st_x"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["xy"][s][:, 1], z["st_y"][s0:s1][m]), (name, t, s)
            assert np.array_equal(st["energy"][s], z["st_e"][s0:s1][m]), (name, t, s, st["energy"][s] - z["st_e"][s0:s1][m])
            assert np.array_equal(st["age"][s], z["st_age"][s0:s1][m]), (name, t, s)
        mp = z["st_s"][s0:s1] == 0
        assert np.array_equal(st["facing"], z["st_face"][s0:s1][mp]), (name, t)
        assert np.array_equal(st["trait"], z["st_trait"][s0:s1][mp]), (name, t)
        assert np.array_equal(st["grass_energy"], z["grass_e"][t]), (name, t)
        assert np.array_equal(sha_f32([o.read_grid(0)]), z["grid_sha"][t]), (name, t)
        assert np.array_equal(st["capture"], z["counters"][t]), (name, t, st["capture"], z["counters"][t])
        assert np.array_equal(st["capture_real"], z["lastp"][t]), (name, t, st["capture_real"], z["lastp"][t])
        if z["all_term"][t] or z["all_trunc"][t]:
            break
    assert int(out["env_status"][0]) == 0
    o.close()
