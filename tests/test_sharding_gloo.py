"""N > 1 path on CPU: world_size-2 gloo job.  Each rank owns a slice of the env instances (stepped here by the
CPU oracle, which stands in for the GPU handle of the rank), no data-path collective; the statistics vector
is all-reduced.  The union of the ranks' results must equal one process stepping all envs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from predpreygrass_b200.config import BASE_CONFIG, N_STATS, STAT_NAMES, make_config
from predpreygrass_b200.sharding import allreduce_stats, max_over_ranks, shard_range

N_ENVS, STEPS = 24, 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rollout(lo, hi):
    """steps envs [lo, hi) of the global job; per-env Philox keys make an env's trajectory independent of its shard"""
    from oracle.oracle import Oracle

    cfg = make_config(BASE_CONFIG, cap_live=(64, 192), seed=99, env_index_base=lo)
    o = Oracle(cfg, hi - lo)
    o.reset()
    rng = np.random.default_rng(5)
    for _ in range(STEPS):
        out = o.outputs()
        a0 = np.full(max(1, out["n"][0]), 4, np.int32)
        a1 = np.full(max(1, out["n"][1]), 4, np.int32)
        # actions keyed by (global env, agent id) so they do not depend on the shard layout
        for s, a in ((0, a0), (1, a1)):
            env = out[f"row_env{s}"].astype(np.int64) + lo
            a[: out["n"][s]] = (env * 7 + out[f"row_agent{s}"] * 3 + _) % 9
        o.step(a0, a1)
    st = o.stats().copy()
    cnt = o.outputs()["env_count"].copy()
    o.close()
    return st, cnt


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N_ENVS, rank, world)
    st, cnt = _rollout(lo, hi)
    t = torch.from_numpy(st.astype(np.int64))
    total = allreduce_stats(t)
    slow = max_over_ranks(1.0 + rank, "cpu")
    q.put((rank, lo, hi, total, cnt, slow))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n, w in ((24, 2), (4096, 8), (10, 3), (5, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_two_rank_gloo_job_equals_single_process():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_stats, ref_cnt = _rollout(0, N_ENVS)
    want = dict(zip(STAT_NAMES, ref_stats.tolist()))
    for rank, lo, hi, total, cnt, slow in res:
        assert slow == 2.0  # max over ranks
        for k in STAT_NAMES:
            if k == "status_envs":
                continue
            assert total[k] == want[k], (k, total[k], want[k])
        assert np.array_equal(cnt, ref_cnt[lo:hi])  # env e of the global job is the same whichever rank owns it
    assert len(res[0][3]) == N_STATS
