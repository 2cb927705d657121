"""Zero-copy consumer path: the compact CUDA row batches go straight into a torch policy and the actions
come straight back — no per-agent dicts, no host copy of observations.

The reference trains with RLlib PPO, one policy per species (`policy_mapping_fn`: agent id prefix -> "predator" /
"prey" policy, eco_evolutionary/tune_ppo.py:42-51), each a `DefaultPPOTorchRLModule` whose encoder is the conv stack
of `eco_evolutionary/utils/networks.py:6-76`.  RLlib's `MultiAgentEnvRunner` would call `env.step(action_dict)` once per
env and build those per-policy batches from dicts; here a step already IS the per-policy batch:

    out.obs[s][:n]   float32 [n, C, R_s, R_s]   every agent of species s of every env, one row each
    env.actions[s]   int32   [cap]              the policy writes row i's action at index i

`DeviceRollout` is the env-runner loop over that layout; `collect()` returns per-policy column dicts with RLlib's
column names (`obs`, `actions`, `rewards`, `terminateds`, `truncateds`, `eps_id`, `agent_index`, `action_logp`) as
CUDA tensors, ready for a learner.  ray is not importable in this image, so the hand-over to an actual
`MultiRLModule` is shown in INTEGRATION.md, not executed here.  torch is the consumer here, not the product: the
environment step stays in the CUDA kernels.
"""
import torch

from .config import VARIANT_STAG

ROW_TERMINATED, ROW_TRUNCATED, ROW_NEWBORN, ROW_FOUNDER = 0x01, 0x02, 0x04, 0x08
STAG_JOIN_SHIFT = 8


class LinearPolicy(torch.nn.Module):
    """the smallest stochastic policy: logits = W · flatten(obs) + b (a stand-in that exercises the plumbing)"""

    def __init__(self, n_in, n_out, device, seed=0):
        super().__init__()
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.lin = torch.nn.Linear(n_in, n_out)
        with torch.no_grad():
            self.lin.weight.copy_(torch.randn(n_out, n_in, generator=g) * 0.05)
            self.lin.bias.zero_()
        self.to(device)

    def forward(self, obs):
        return self.lin(obs.flatten(1))


class ConvPolicy(torch.nn.Module):
    """the reference's module shape (networks.py:28-50): L = (R - 1) // 2 conv layers 3x3 stride 1 with 16, 32, 64, 64...
    filters, then fully connected [256, 256] (or [384, 256] for > 20 actions), ReLU, a logits head"""

    def __init__(self, channels, window, n_out, device):
        super().__init__()
        L = (window - 1) // 2
        widths = ([16, 32, 64] + [64] * max(0, L - 3))[:L]
        layers, c = [], channels
        for w in widths:
            layers += [torch.nn.Conv2d(c, w, 3, stride=1, padding=1), torch.nn.ReLU()]  # RLlib pads "same"
            c = w
        self.conv = torch.nn.Sequential(*layers)
        hid = [384, 256] if n_out > 20 else [256, 256]
        self.fc = torch.nn.Sequential(torch.nn.Linear(c * window * window, hid[0]), torch.nn.ReLU(), torch.nn.Linear(hid[0], hid[1]),
                                      torch.nn.ReLU(), torch.nn.Linear(hid[1], n_out))
        self.to(device)

    def forward(self, obs):
        return self.fc(self.conv(obs).flatten(1))


class DeviceRollout:
    """env-runner loop on the device: policy(obs rows) -> actions in place -> ppg_step.

    policies: (predator_module, prey_module); a module maps float32 [n, C, R, R] to logits [n, n_logits].  For the STAG
    predator (`MultiDiscrete([n_moves, 2])`, STAG:1813-1816) the logits are n_moves + 2 wide: move logits, then join_hunt logits.
    """

    def __init__(self, env, policies, stream=None, sample=True, seed=0):
        self.env, self.policies, self.stream, self.sample = env, policies, stream, sample
        self.gen = torch.Generator(device=env.device).manual_seed(seed)
        self.variant = env.cfg.variant
        self.last = None

    def _ctx(self):
        return torch.cuda.stream(self.stream) if self.stream is not None else torch.cuda.stream(torch.cuda.current_stream(self.env.device))

    def _pick(self, logits):
        if not self.sample:
            return logits.argmax(-1), None
        logp = torch.log_softmax(logits, -1)
        a = torch.multinomial(logp.exp(), 1, generator=self.gen).squeeze(1)
        return a, logp.gather(1, a.unsqueeze(1)).squeeze(1)

    @torch.no_grad()
    def act(self):
        """policy forward on the rows of the last output; actions written into env.actions in place.
        -> per species (n, actions int64 [n], logp or None)"""
        env, out = self.env, self.env.out
        n = out.counts()  # 16 bytes device -> host: the row counts of the last output
        res = []
        for s in range(2):
            k = n[s]
            if k == 0:
                res.append((0, None, None))
                continue
            logits = self.policies[s](out.obs[s][:k])
            if self.variant == VARIANT_STAG and s == 0:
                nm = env.n_actions(0)
                mv, lp1 = self._pick(logits[:, :nm])
                jn, lp2 = self._pick(logits[:, nm:nm + 2])
                a = mv | (jn << STAG_JOIN_SHIFT)
                lp = None if lp1 is None else lp1 + lp2
            else:
                a, lp = self._pick(logits)
            env.actions[s][:k] = a.to(torch.int32)
            res.append((k, a, lp))
        return res

    @torch.no_grad()
    def step(self):
        with self._ctx():
            self.last = self.act()
            return self.env.step()

    @torch.no_grad()
    def collect(self, n_steps, check=False):
        """n_steps of experience -> {"predator": columns, "prey": columns}.  A transition pairs the observation and action of
        an agent in output t with the reward / termination flags output t+1 reports for the same (env, agent id).  The rows of
        t+1 are grouped by env with last step's newborns merged in, so the pairing is a key match on the device (sort +
        gather), not a positional one.  Rows produced by reset() (ROW_FOUNDER) or born in t+1 (ROW_NEWBORN) have no
        predecessor; rows already terminated / truncated in t have no successor."""
        names = ("predator", "prey")
        cols = {nm: {k: [] for k in ("obs", "actions", "action_logp", "rewards", "terminateds", "truncateds", "eps_id", "agent_index")} for nm in names}
        npos = (int(self.env.cfg.n_possible[0]), int(self.env.cfg.n_possible[1]))
        with self._ctx():
            for _ in range(n_steps):
                prev = self.env.out
                acted = self.act()
                keep = []
                for s in range(2):
                    k, a, lp = acted[s]
                    if k == 0:
                        keep.append(None)
                        continue
                    live = (prev.flags[s][:k] & (ROW_TERMINATED | ROW_TRUNCATED)) == 0
                    key = prev.row_env[s][:k].to(torch.int64) * npos[s] + prev.row_agent[s][:k].to(torch.int64)
                    key, order = torch.sort(key[live])
                    keep.append((prev.obs[s][:k][live][order], a[live][order], None if lp is None else lp[live][order], key))
                out = self.env.step()
                n_old = out.n_rows[:2].tolist()
                for s, nm in enumerate(names):
                    if keep[s] is None:
                        continue
                    obs, a, lp, key = keep[s]
                    k2 = n_old[s]
                    f = out.flags[s][:k2]
                    succ = (f & ROW_FOUNDER) == 0
                    key2 = out.row_env[s][:k2].to(torch.int64) * npos[s] + out.row_agent[s][:k2].to(torch.int64)
                    key2, order2 = torch.sort(key2[succ])
                    if check:
                        assert key2.shape == key.shape and bool((key2 == key).all()), "row pairing broken"
                    c = cols[nm]
                    c["obs"].append(obs)
                    c["actions"].append(a)
                    if lp is not None:
                        c["action_logp"].append(lp)
                    c["rewards"].append(out.reward[s][:k2][succ][order2])
                    f2 = f[succ][order2]
                    c["terminateds"].append((f2 & ROW_TERMINATED) != 0)
                    c["truncateds"].append((f2 & ROW_TRUNCATED) != 0)
                    c["eps_id"].append(key // npos[s])
                    c["agent_index"].append(key % npos[s])
        return {nm: {k: (torch.cat(v) if v else None) for k, v in c.items()} for nm, c in cols.items()}
