"""connector.DeviceRollout (the compact CUDA row batches straight into a torch policy, the actions straight back):
the transitions `collect()` hands to a learner are checked by an independent route — a twin handle of the same config and
seed replays the collected (env, agent, action) triples step by step and must report exactly the rewards and termination flags
the collected transitions carry, for the observations they carry."""
import pytest

pytestmark = pytest.mark.gpu

ROW_ENDED, ROW_FOUNDER = 0x03, 0x08


def _cfg(variant):
    from predpreygrass_b200.config import BASE_CONFIG, ECO_CONFIG, STAG_CONFIG, VARIANT_ECO, VARIANT_STAG, make_config

    if variant == "base":
        return make_config(dict(BASE_CONFIG, max_steps=25), reward_mode="additive", cap_live=(32, 128), seed=3)
    if variant == "eco":
        return make_config(dict(ECO_CONFIG, max_steps=30, energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0,
                                predator_creation_energy_threshold=8.0), variant=VARIANT_ECO, cap_live=(64, 128), seed=3)
    return make_config(dict(STAG_CONFIG, max_steps=30), variant=VARIANT_STAG, cap_live=(64, 192), seed=3)


@pytest.mark.parametrize("variant", ["base", "eco", "stag"])
def test_collected_transitions_replay_on_a_twin_handle(variant):
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass
    from predpreygrass_b200.connector import DeviceRollout, LinearPolicy

    cfg, B = _cfg(variant), 96
    a, b = BatchedPredPreyGrass(cfg, B), BatchedPredPreyGrass(cfg, B)
    try:
        a.reset()
        b.reset()
        n_logits = [a.n_actions(0) + (2 if variant == "stag" else 0), a.n_actions(1)]  # STAG predator: move logits + join_hunt logits
        pols = [LinearPolicy(a.C * a.R[s] ** 2, n_logits[s], device=a.device, seed=11 + s) for s in range(2)]
        ro = DeviceRollout(a, pols, sample=True, seed=5)
        npos = (int(cfg.n_possible[0]), int(cfg.n_possible[1]))
        names = ("predator", "prey")
        n_tr = [0, 0]
        ended = 0
        for t in range(45):
            before = b.out
            nb = before.counts()
            prev_obs = [before.obs[s][: nb[s]].clone() for s in range(2)]
            prev_key = [before.row_env[s][: nb[s]].to(torch.int64) * npos[s] + before.row_agent[s][: nb[s]].to(torch.int64) for s in range(2)]
            prev_live = [(before.flags[s][: nb[s]] & ROW_ENDED) == 0 for s in range(2)]
            cols = ro.collect(1, check=True)  # one step of the rollout handle: policy on its rows, actions in place, ppg_step
            for s, nm in enumerate(names):
                c = cols[nm]
                if c["obs"] is None or c["obs"].shape[0] == 0:  # nobody acts (the step after every env's time limit)
                    assert int(prev_live[s].sum()) == 0
                    continue
                key_c = c["eps_id"] * npos[s] + c["agent_index"]
                assert bool((key_c[1:] > key_c[:-1]).all())  # one transition per acting agent, sorted by (env, agent id)
                rows = torch.nonzero(prev_live[s]).squeeze(1)
                assert rows.numel() == key_c.numel(), (variant, t, nm)
                idx = torch.searchsorted(key_c, prev_key[s][rows])
                assert bool((key_c[idx] == prev_key[s][rows]).all()), (variant, t, nm)  # every live row of the twin has its transition
                assert torch.equal(c["obs"][idx], prev_obs[s][rows]), (variant, t, nm)  # ... which carries that row's observation
                mv, jn = c["actions"] & 0xFF, c["actions"] >> 8  # STAG predators: move | join_hunt << 8
                assert int(mv.min()) >= 0 and int(mv.max()) < a.n_actions(s) and 0 <= int(jn.min()) and int(jn.max()) <= (1 if variant == "stag" and s == 0 else 0)
                assert c["action_logp"].shape == c["actions"].shape and bool((c["action_logp"] <= 0).all())
                b.actions[s].zero_()
                b.actions[s][rows] = c["actions"][idx].to(torch.int32)
                n_tr[s] += int(key_c.numel())
            out = b.step()
            n_old = out.n_rows[:2].tolist()
            for s, nm in enumerate(names):
                c = cols[nm]
                if c["obs"] is None or c["obs"].shape[0] == 0:
                    continue
                k2 = n_old[s]
                f = out.flags[s][:k2]
                succ = (f & ROW_FOUNDER) == 0
                key2 = out.row_env[s][:k2].to(torch.int64) * npos[s] + out.row_agent[s][:k2].to(torch.int64)
                key2, order = torch.sort(key2[succ])
                assert torch.equal(key2, c["eps_id"] * npos[s] + c["agent_index"]), (variant, t, nm)
                assert torch.equal(out.reward[s][:k2][succ][order], c["rewards"]), (variant, t, nm)
                f2 = f[succ][order]
                assert torch.equal((f2 & 0x01) != 0, c["terminateds"]) and torch.equal((f2 & 0x02) != 0, c["truncateds"]), (variant, t, nm)
                ended += int(((f2 & ROW_ENDED) != 0).sum())
            # both handles are in the same state again
            assert torch.equal(a.out.n_rows, b.out.n_rows) and torch.equal(a.out.env_step, b.out.env_step)
        assert n_tr[0] > 1000 and n_tr[1] > 1000 and ended > 100, (n_tr, ended)  # deaths and time limits were among the transitions
        sa, sb = a.stats(), b.stats()
        assert sa["episodes"] >= B and all(sa[k] == sb[k] for k in ("episodes", "env_steps", "agent_steps", "births_pred", "births_prey"))
    finally:
        a.close()
        b.close()


def test_conv_policy_shape_and_in_place_actions():
    """the reference's module shape (networks.py:28-50) on the row batch: logits per row, actions written in place"""
    import torch

    from predpreygrass_b200.batched import BatchedPredPreyGrass
    from predpreygrass_b200.connector import ConvPolicy, DeviceRollout

    cfg = _cfg("eco")
    env = BatchedPredPreyGrass(cfg, 32)
    try:
        env.reset()
        pols = [ConvPolicy(env.C, env.R[s], env.n_actions(s), env.device) for s in range(2)]
        ro = DeviceRollout(env, pols, sample=False)
        for _ in range(3):
            ro.step()
        (k0, a0, _), (k1, a1, _) = ro.last
        n = env.out.counts()
        assert k0 > 0 and k1 > 0 and a0.shape == (k0,) and a1.shape == (k1,) and n[0] > 0 and n[1] > 0
        assert int(a0.max()) < env.n_actions(0) and int(a1.max()) < env.n_actions(1)
        assert torch.equal(env.actions[0][:k0], a0.to(torch.int32)) and torch.equal(env.actions[1][:k1], a1.to(torch.int32))
    finally:
        env.close()
