"""Side-by-side run of the CUDA path and the CPU oracle on identical inputs (test infrastructure).

Used by the `-m gpu` tests, by __graft_entry__.smoke() and by bench.py's correctness check."""
import numpy as np


EXACT_KEYS = ("row_env", "row_agent", "flags", "old_off", "new_off", "new_cnt")


def compare_outputs(g, o, where="", obs_rtol=0.0, reward_rtol=0.0):
    """g: BatchedPredPreyGrass.outputs_numpy(), o: Oracle.outputs().  Integers, ids, flags and row
    layout bit-exact; rewards/observations exact by default (both sides round the same float64
    value to float32), or within the given relative tolerance (north_star: 1e-5)."""
    assert g["n_old"] == o["n_old"] and g["n_new"] == o["n_new"], (where, g["n_old"], o["n_old"], g["n_new"], o["n_new"])
    for s in range(2):
        for k in EXACT_KEYS:
            a, b = g[f"{k}{s}"], o[f"{k}{s}"]
            if not np.array_equal(a, b):
                bad = np.nonzero(a != b)[0][:5]
                raise AssertionError(f"{where}: {k}{s} differs at {bad}: gpu={a[bad]} oracle={b[bad]}")
        for k, tol in (("reward", reward_rtol), ("obs", obs_rtol)):
            a, b = g[f"{k}{s}"], o[f"{k}{s}"]
            ok = np.array_equal(a, b) if tol == 0.0 else np.allclose(a, b, rtol=tol, atol=0.0)
            if not ok:
                bad = np.argwhere(a != b)[:5]
                rows = [int(r[0]) for r in bad]
                raise AssertionError(
                    f"{where}: {k}{s} differs at {bad.tolist()} (env {g[f'row_env{s}'][rows]}, agent {g[f'row_agent{s}'][rows]}, "
                    f"flags {g[f'flags{s}'][rows]}): gpu={a[tuple(bad.T)]} oracle={b[tuple(bad.T)]}")
        if tol != 0.0:
            # occupancy (which cells are non-zero) must be bit-exact even under a value tolerance
            assert np.array_equal(g[f"obs{s}"] != 0, o[f"obs{s}"] != 0), f"{where}: obs{s} occupancy differs"
    for k in ("env_flags", "env_status", "env_step", "env_count"):
        if not np.array_equal(g[k], o[k]):
            bad = np.argwhere(g[k] != o[k])[:5]
            raise AssertionError(f"{where}: {k} differs at {bad.tolist()}: gpu={g[k][tuple(bad.T)]} oracle={o[k][tuple(bad.T)]}")


def compare_env_state(gpu, oracle, env, where=""):
    eco, stag = gpu.cfg.variant == 1, gpu.cfg.variant == 2
    if stag:
        a, b = gpu.read_env_stag(env), oracle.read_env_stag(env)
        for s in range(2):
            assert np.array_equal(a["age"][s], b["age"][s]), (where, env, s, a["age"][s], b["age"][s])
        assert np.array_equal(a["facing"], b["facing"]), (where, env, a["facing"], b["facing"])
        assert np.array_equal(a["trait"], b["trait"]), (where, env, a["trait"], b["trait"])
        assert np.array_equal(a["capture"], b["capture"]), (where, env, a["capture"], b["capture"])
        assert np.array_equal(a["capture_real"], b["capture_real"]), (where, env, a["capture_real"], b["capture_real"])
    else:
        a, b = (gpu.read_env_eco(env), oracle.read_env_eco(env)) if eco else (gpu.read_env(env), oracle.read_env(env))
    if eco:
        for s in range(2):
            assert np.array_equal(a["age"][s], b["age"][s]), (where, env, s, a["age"][s], b["age"][s])
            assert np.array_equal(a["speed"][s], b["speed"][s]), (where, env, s)
        assert np.array_equal(a["dead_prey"], b["dead_prey"]), (where, env)
        assert np.array_equal(a["active_num"], b["active_num"]), (where, env)
        if gpu.cfg.track_episode_sums:  # per-episode totals and the trait variants' event counters (float64, accumulated in order)
            a_ep, b_ep = gpu.read_episode_eco(env), oracle.read_episode_eco(env)
            assert a_ep["spawned"] == b_ep["spawned"], (where, env, a_ep, b_ep)
            for k in ("distance", "move_energy"):  # the device adds per-lane partial sums: equal to the last few ulps
                assert np.allclose(a_ep[k], b_ep[k], rtol=1e-12, atol=0.0), (where, env, k, a_ep, b_ep)
            if gpu.cfg.trait_mode != 0:
                ge, oe = gpu.read_episode_events_eco(env), oracle.read_episode_events_eco(env)
                assert ge == oe, (where, env, ge, oe)
        if gpu.cfg.trait_mode == 4:  # cadence: the move accumulators (float64, bit-exact)
            ga, oa = gpu.read_env_acc(env), oracle.read_env_acc(env)
            for s in range(2):
                assert np.array_equal(ga[s], oa[s]), (where, env, s, ga[s], oa[s])
    for s in range(2):
        assert np.array_equal(a["ids"][s], b["ids"][s]), (where, env, s, a["ids"][s], b["ids"][s])
        assert np.array_equal(a["xy"][s], b["xy"][s]), (where, env, s)
        assert np.array_equal(a["energy"][s], b["energy"][s]), (where, env, s, a["energy"][s], b["energy"][s])
    assert np.array_equal(a["grass_xy"], b["grass_xy"]), (where, env)
    assert np.array_equal(a["grass_energy"], b["grass_energy"]), (where, env)


def lockstep_parity(cfg, n_envs, steps, *, tape=None, reals=None, seeds=None, action_seed=1234, threads=8, state_envs=(0,),
                    check_every=1, device=0):
    """Run both sides for `steps` steps with identical Philox-keyed random actions; compare every
    `check_every` steps (always the last).  Returns the oracle's stats dict for reporting."""
    import torch

    from oracle.oracle import Oracle
    from predpreygrass_b200.batched import BatchedPredPreyGrass
    from predpreygrass_b200.config import STAT_NAMES

    gpu = BatchedPredPreyGrass(cfg, n_envs, device=device)
    ora = Oracle(cfg, n_envs, threads=threads)
    try:
        if tape is not None:
            gpu.load_tape(tape, reals)
            ora.load_tape(tape, reals)
        gpu.reset(seeds)
        ora.reset(seeds)
        compare_outputs(gpu.outputs_numpy(), ora.outputs(), "reset")
        for t in range(steps):
            a0, a1 = gpu.random_actions(action_seed)
            o0, o1 = ora.random_actions(action_seed)
            if t % check_every == 0:
                n0, n1 = len(o0), len(o1)
                assert np.array_equal(a0[:n0].cpu().numpy(), o0) and np.array_equal(a1[:n1].cpu().numpy(), o1), f"actions differ at step {t}"
            gpu.step(a0, a1)
            A0 = np.zeros(max(1, len(o0)), np.int32); A0[: len(o0)] = o0
            A1 = np.zeros(max(1, len(o1)), np.int32); A1[: len(o1)] = o1
            ora.step(A0, A1)
            if t % check_every == 0 or t == steps - 1:
                compare_outputs(gpu.outputs_numpy(), ora.outputs(), f"step {t}")
                for e in state_envs:
                    if not (ora.outputs()["env_flags"][e] & 3):  # state of a finished episode is not kept
                        compare_env_state(gpu, ora, e, f"step {t}")
        gs, os_ = gpu.stats(), dict(zip(STAT_NAMES, ora.stats().tolist()))
        for k in STAT_NAMES:
            if k in ("rows_pred", "rows_prey", "reserved"):
                continue
            assert gs[k] == os_[k], (k, gs[k], os_[k])
        torch.cuda.synchronize()
        gs["status_or"] = int(np.bitwise_or.reduce(ora.outputs()["env_status"]))
        return gs
    finally:
        gpu.close()
        ora.close()
