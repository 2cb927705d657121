"""The path `bench.py` times — `ppg_rollout_random` (device-side random actions + step, issued from C), env groups on their own
CUDA streams (`pipelined.PipelinedPredPreyGrass`), K steps queued without any host synchronisation — against the
oracle: after K steps every output array of every group equals the oracle's, stepped one call at a time with the same
Philox-keyed actions.  The measured path is the parity-tested path."""
import numpy as np
import pytest

from predpreygrass_b200.config import (BASE_CONFIG, ECO_CONFIG, STAG_CONFIG, STAT_NAMES, VARIANT_BASE, VARIANT_ECO, VARIANT_STAG,
                                       make_config)
from tests.parity import compare_env_state, compare_outputs

pytestmark = pytest.mark.gpu


def _factory(variant):
    if variant == "base":
        return lambda off: make_config(dict(BASE_CONFIG, max_steps=40), variant=VARIANT_BASE, reward_mode="additive", cap_live=(32, 128), seed=9,
                                       env_index_base=off)
    if variant == "eco":
        return lambda off: make_config(dict(ECO_CONFIG, max_steps=40), variant=VARIANT_ECO, cap_live=(32, 96), seed=9, env_index_base=off)
    return lambda off: make_config(dict(STAG_CONFIG, max_steps=40), variant=VARIANT_STAG, cap_live=(64, 192), seed=9, env_index_base=off)


# (variant, env groups, launch chain, steps per call).  History of the launch chain (ppg_set_pdl_chain): this test FAILED for every
# variant with the chain on and the 55 steps queued in one call (row counts drifting from the oracle's after a few steps) while
# the per-step parity suite, which synchronises after every step, was green.  Cause (SASS): in the action kernels the compiler had
# hoisted the load of n_rows[0] — a `const __restrict__` pointer — above griddepcontrol.wait, so the actions of step t+1 covered
# the row count of step t-1 whenever the previous step's kernels were still running.  Fixed (volatile load after the wait); the
# chain-on cases pass since and are part of the suite.  The chain stays off by default (include/ppg.h).
CASES = [(v, g, 0, c) for v in ("base", "eco", "stag") for g in (1, 2) for c in (55, 1) if not (g == 2 and c == 1)]
CASES += [(v, 1, 1, c) for v in ("base", "eco", "stag") for c in (1, 55)]


@pytest.mark.parametrize("variant,groups,pdl,chunk", CASES)
def test_rollout_random_on_streams_matches_the_oracle(variant, groups, pdl, chunk):
    import torch

    from oracle.oracle import Oracle
    from predpreygrass_b200 import _lib
    from predpreygrass_b200.pipelined import PipelinedPredPreyGrass

    B, K, seed = 768, 55, 4321
    factory = _factory(variant)
    L = _lib.load()
    pipe = PipelinedPredPreyGrass(factory, B, groups=groups)
    before = L.ppg_set_pdl_chain(pdl)  # process-wide switch
    try:
        pipe.reset()
        for _ in range(K // chunk):  # chunk = K: one call, no host synchronisation for K steps (what bench.py times); chunk = 1: a sync per step
            pipe.rollout_random(chunk, seed)
            torch.cuda.synchronize()
        per = B // groups
        for g, e in enumerate(pipe.envs):
            ora = Oracle(factory(g * per), per, threads=8)
            try:
                ora.reset(None)
                for _ in range(K):
                    o0, o1 = ora.random_actions(seed)
                    a0 = np.zeros(max(1, len(o0)), np.int32); a0[: len(o0)] = o0
                    a1 = np.zeros(max(1, len(o1)), np.int32); a1[: len(o1)] = o1
                    ora.step(a0, a1)
                want = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in ora.outputs().items()}  # copies: a failure report must not touch freed memory
                compare_outputs(e.outputs_numpy(), want, f"{variant} group {g}")
                for env in (0, per // 2, per - 1):
                    if not (ora.outputs()["env_flags"][env] & 3):
                        compare_env_state(e, ora, env, f"{variant} group {g}")
                gs, os_ = e.stats(), dict(zip(STAT_NAMES, ora.stats().tolist()))
                for k in STAT_NAMES:
                    if k not in ("rows_pred", "rows_prey", "reserved"):
                        assert gs[k] == os_[k], (variant, g, k, gs[k], os_[k])
                assert gs["episodes"] >= per and gs["status_envs"] == 0
            finally:
                ora.close()
    finally:
        L.ppg_set_pdl_chain(0)  # the default
        pipe.close()
