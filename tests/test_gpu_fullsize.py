"""The BASELINE sizes (`-m gpu`): ADD 16 384 envs (configs[2]), ECO 16 384 envs (configs[3]) and one GPU's 8192-env slice
of STAG's 65 536 (configs[4]).

1. Oracle lockstep AT those sizes: 100 steps with every output array compared bit for bit every 10 steps (and at the
   end), state of sampled envs included — the oracle does ~1.5e6 agent-steps/s on the box's host cores, i.e. 20-40 s
   per config.
2. Size-independent properties: determinism across handles, sharding invariance (two handles with `env_index_base` =
   one handle with all envs: the multi-GPU contract of DESIGN.md §6), row bookkeeping and the own-cell observation."""
import os

import pytest

from predpreygrass_b200.config import ECO_CONFIG, STAG_CONFIG, VARIANT_ECO, VARIANT_STAG, make_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["add", "eco", "stag"])
def test_fullsize_oracle_lockstep(which):
    from predpreygrass_b200.config import BASE_CONFIG
    from tests.parity import lockstep_parity

    if which == "add":
        cfg, B = make_config(BASE_CONFIG, reward_mode="additive", cap_live=(32, 128), seed=41), 16384
    elif which == "eco":
        cfg, B = make_config(ECO_CONFIG, variant=VARIANT_ECO, cap_live=(32, 96), seed=42), 16384
    else:
        # slot capacities 64 + 192: a few envs fill their prey list within 100 steps — the oracle suppresses the same births
        # and raises the same status bit, so the comparison covers that path too
        cfg, B = make_config(STAG_CONFIG, variant=VARIANT_STAG, cap_live=(64, 192), seed=43), 8192
    st = lockstep_parity(cfg, B, 100, state_envs=(0, B // 2, B - 1), check_every=10, threads=os.cpu_count() or 8)
    assert st["env_steps"] >= 95 * B and st["births_prey"] > 0


def _check_rows(h, cfg, B, own_channel):
    import torch

    o = h.out
    n = o.n_rows.tolist()
    for s in range(2):
        k = n[s] + n[2 + s]
        re = o.row_env[s][: n[s]]
        assert bool((re[1:] >= re[:-1]).all())  # rows of the agents that acted: grouped by env, ascending
        ended = (o.flags[s][:k] & 3) != 0
        R = cfg.obs_range[s]
        if cfg.variant == VARIANT_STAG and s == 0:
            continue  # predators look through a window shifted along their facing: the centre is not their own cell
        centre = o.obs[s][:k, own_channel(s), R // 2, R // 2]
        if cfg.variant == VARIANT_STAG:
            centre = centre + o.obs[s][:k, own_channel(s) + 1, R // 2, R // 2]  # mammoths on channel 2, rabbits on channel 3
        assert float((centre[~ended] > 0).float().mean()) > 0.995  # own energy at the window centre
        live = (~ended).long()
        per_env = torch.zeros(B, dtype=torch.long, device=live.device).index_add_(0, o.row_env[s][:k].long(), live)
        running = (o.env_flags & 7) == 0
        if cfg.variant == VARIANT_STAG:
            assert torch.equal(per_env[running], o.env_count.long()[running, s])
    if cfg.variant == VARIANT_STAG:  # ended agents are observed as all-zero rows (STAG:596-612)
        for s in range(2):
            k = n[s] + n[2 + s]
            ended = (o.flags[s][:k] & 1) != 0
            if bool(ended.any()):
                assert float(o.obs[s][:k][ended].abs().sum()) == 0.0


@pytest.mark.parametrize("which", ["eco", "stag"])
def test_fullsize_determinism_sharding_and_rows(which):
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass

    if which == "eco":
        B, caps, src, var = 16384, (32, 96), ECO_CONFIG, VARIANT_ECO
        own = lambda s: s  # noqa: E731  channels: predators, prey, grass (+ speed plane)
    else:
        B, caps, src, var = 8192, (160, 640), STAG_CONFIG, VARIANT_STAG
        own = lambda s: 1 + s  # noqa: E731  walls, predators, mammoths, rabbits, grass
    cfg = make_config(src, variant=var, cap_live=caps, seed=31)
    whole = BatchedPredPreyGrass(cfg, B)
    half = [BatchedPredPreyGrass(make_config(src, variant=var, cap_live=caps, seed=31, env_index_base=k * (B // 2)), B // 2) for k in range(2)]
    for h in [whole] + half:
        h.reset()
    steps = 60 if which == "eco" else 40
    for t in range(steps):
        for h in [whole] + half:
            a0, a1 = h.random_actions(7)
            h.step(a0, a1)
        if t % 20 == 19 or t == steps - 1:
            n = whole.out.n_rows.tolist()
            nh = [h.out.n_rows.tolist() for h in half]
            assert [n[i] for i in range(4)] == [nh[0][i] + nh[1][i] for i in range(4)]
            for s in range(2):
                # rows of the agents that acted: shard 0's, then shard 1's; newborn rows likewise
                parts = [half[0].out.obs[s][: nh[0][s]], half[1].out.obs[s][: nh[1][s]],
                         half[0].out.obs[s][nh[0][s]: nh[0][s] + nh[0][2 + s]], half[1].out.obs[s][nh[1][s]: nh[1][s] + nh[1][2 + s]]]
                assert torch.equal(whole.out.obs[s][: n[s] + n[2 + s]], torch.cat(parts))
                ids = [half[0].out.row_agent[s][: nh[0][s]], half[1].out.row_agent[s][: nh[1][s]],
                       half[0].out.row_agent[s][nh[0][s]: nh[0][s] + nh[0][2 + s]], half[1].out.row_agent[s][nh[1][s]: nh[1][s] + nh[1][2 + s]]]
                assert torch.equal(whole.out.row_agent[s][: n[s] + n[2 + s]], torch.cat(ids))
                rw = [half[0].out.reward[s][: nh[0][s]], half[1].out.reward[s][: nh[1][s]],
                      half[0].out.reward[s][nh[0][s]: nh[0][s] + nh[0][2 + s]], half[1].out.reward[s][nh[1][s]: nh[1][s] + nh[1][2 + s]]]
                assert torch.equal(whole.out.reward[s][: n[s] + n[2 + s]], torch.cat(rw))
            _check_rows(whole, cfg, B, own)
    st = whole.stats()
    sh = [h.stats() for h in half]
    for k in ("env_steps", "agent_steps", "episodes", "births_pred", "births_prey", "eaten_prey", "grass_eaten", "capture_attempts"):
        assert st[k] == sh[0][k] + sh[1][k], k  # what the all-reduce of the statistics adds up
    assert st["status_envs"] == 0
    for h in [whole] + half:
        h.close()
