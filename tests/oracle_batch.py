"""CPU stand-in for `BatchedPredPreyGrass` over the C oracle: lets the dict adapters' HOST logic (row <-> dict mapping, the
event recorder, metrics) run in the `-m "not gpu"` suite.  Test infrastructure only — the product never imports `oracle/`."""
import numpy as np

from oracle.oracle import Oracle


class OracleBatch:
    """the part of the BatchedPredPreyGrass surface the 1-env dict adapters use"""

    def __init__(self, cfg, n_envs, device=0):
        assert n_envs == 1
        self.cfg, self.n_envs, self.device = cfg, n_envs, "cpu"
        self.o = Oracle(cfg, n_envs)
        self.row_capacity = (max(cfg.n_possible[0], cfg.cap_live[0], 1), max(cfg.n_possible[1], cfg.cap_live[1], 1))

    def close(self):
        self.o.close()

    def load_tape(self, cells, reals=None):
        self.o.load_tape(cells, reals)

    def reset(self, seeds=None, mask=None):
        self.o.reset(seeds, mask)

    def outputs_numpy(self):
        out = self.o.outputs()
        # The oracle's literal single-env interface follows BASE's own protocol (the step after max_steps real steps is the
        # truncation call, BASE:228-238); the batch interface of include/ppg.h flags the env on the max_steps-th step
        # itself (PPG_ENV_TRUNCATED) and the adapter plays the extra call.  Bridge the two for the BASE family.
        if self.cfg.variant == 0 and self.cfg.max_steps > 0 and int(out["env_step"][0]) >= self.cfg.max_steps and not out["env_flags"][0] & 1:
            out["env_flags"][0] |= 2
        return out

    def step_ordered(self, a0, a1, ord0, ord1):
        out = self.o.outputs()
        acts, order = (a0.numpy(), a1.numpy()), (ord0.numpy(), ord1.numpy())
        sp, ids, av = [], [], []
        for s in range(2):
            rows = list(range(int(out[f"old_off{s}"][0]), int(out[f"old_off{s}"][1])))
            r0 = int(out[f"new_off{s}"][0])
            rows += list(range(r0, r0 + int(out[f"new_cnt{s}"][0])))
            rows = [r for r in rows if not out[f"flags{s}"][r] & 3]  # live agents only
            for r in sorted(rows, key=lambda r: order[s][r]):
                sp.append(s); ids.append(int(out[f"row_agent{s}"][r])); av.append(int(acts[s][r]))
        self.o.env_step_ordered(0, sp, ids, av)

    def read_env_eco(self, env):
        return self.o.read_env_eco(env)

    def read_episode_eco(self, env):
        return self.o.read_episode_eco(env)

    def read_episode_events_eco(self, env):
        return self.o.read_episode_events_eco(env)

    def read_env(self, env):
        return self.o.read_env(env)

    def read_env_stag(self, env):
        return self.o.read_env_stag(env)

    def read_env_acc(self, env):
        return self.o.read_env_acc(env)
