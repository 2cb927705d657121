"""PredPreyGrass — the reference's RLlib `MultiAgentEnv` interface over the CUDA step.

Drop-in for `PredPreyGrass(MultiAgentEnv)` of the reference's BASE family
(BASE = predpreygrass/non_evolutionary/base_environment/predpreygrass_rllib_env.py and the
project_reward_shaping variants): same constructor (`config` dict, same keys and defaults,
BASE:18-61), same `reset(*, seed, options) -> (obs, infos)` / `step(action_dict) -> (obs, rewards,
terminations, truncations, infos)` with agent-id string dicts and `"__all__"` (BASE:129,219,456-473),
same attributes (`agents`, `possible_agents`, `observation_spaces`, `action_spaces`,
`agent_positions`, `agent_energies`, `grass_positions`, `grass_energies`, `agents_just_ate`,
`current_step`, BASE:70-127) and `get_state_snapshot()/restore_state_snapshot()` (BASE:768-804).

One object = one environment instance of a `BatchedPredPreyGrass` handle (by default its own
1-env handle).  All simulation work is done by the CUDA kernels; this file only converts between
the handle's row batch and the reference's dicts.  For throughput use `BatchedPredPreyGrass`
directly (thousands of envs, tensors stay on the GPU); this adapter exists so that the reference's
scripts (`random_policy.py`, `evaluate_*.py`, `tune_ppo_*.py:env_creator`) run unchanged.

Differences from the reference, all deliberate:
  * observations are computed in fp32 on the device and returned as float64 arrays (the declared
    Box dtype, BASE:88-94): values equal the reference's to fp32 rounding (<= 6e-8 relative);
  * `reset(seed=s)` reproduces the reference's initial placement exactly (numpy PCG64 +
    CPython set order, BASE:156-187, evaluated on the host and handed to the device as a replay
    tape); the rare spawn-fallback draw (BASE:760-764, global `np.random`) comes from the
    device's Philox stream instead;
  * the movement order of a step is the iteration order of `action_dict` (BASE:259), as in the
    reference; it is passed to the device as a per-row rank (`ppg_step_ordered`).
"""
import numpy as np

from .config import (ENV_TERMINATED, ENV_TRUNCATED, REWARD_MODES, ROW_ATE, ROW_TERMINATED, make_config)

try:  # the real base class when ray is installed, so RLlib accepts the object
    from ray.rllib.env.multi_agent_env import MultiAgentEnv as _Base
except Exception:  # noqa: BLE001
    class _Base:  # minimal stand-in with the same no-op surface
        def __init__(self):
            pass

        def reset(self, *, seed=None, options=None):
            pass

        def close(self):
            pass

try:
    from gymnasium.spaces import Box, Discrete
except Exception:  # noqa: BLE001
    class Box:  # the attributes the reference's scripts read (tune_ppo_base_environment.py:92-104)
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)
            self._rng = np.random.default_rng()

        def sample(self):
            return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

    class Discrete:
        def __init__(self, n):
            self.n = int(n)
            self._rng = np.random.default_rng()

        def sample(self):
            return int(self._rng.integers(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

_SPECIES = ("predator", "prey")

# include/ppg.h PPG_STATUS_*
_ST_SLOT_OVERFLOW, _ST_NO_SPAWN_CELL, _ST_TAPE_EXHAUSTED, _ST_BAD_ACTION, _ST_ID_POOL_EMPTY = 0x01, 0x02, 0x04, 0x08, 0x10


def raise_on_status(status, variant=0):
    """The in-kernel conditions the reference raises on (sticky per-env status bits, include/ppg.h) become the
    reference's exceptions in the dict adapters: a drop-in user must not get a silently diverging trajectory.
    (PPG_STATUS_TAPE_EXHAUSTED is not an error here: the adapters tape only the reset draws and let the device's
    Philox stream continue.)"""
    if status & _ST_BAD_ACTION:
        raise KeyError("action outside the action space (BASE:502 indexes action_to_move_tuple)")
    if status & _ST_ID_POOL_EMPTY:
        # ECO:1104-1111 prints the message and raises SystemExit(msg)
        raise SystemExit("No available agent IDs left: increase n_possible_predators / n_possible_prey")
    if status & _ST_NO_SPAWN_CELL:
        raise RuntimeError("no free cell for a newborn (BASE:766 / ECO:1142-1143)")
    if status & _ST_SLOT_OVERFLOW:
        raise RuntimeError("device slot capacity exhausted: a birth the reference allows was suppressed; raise config['cap_live']")


def _split(agent):
    kind, idx = agent.rsplit("_", 1)
    return (0 if kind == "predator" else 1), int(idx)


def reference_initial_cells(seed, grid_size, n_total):
    """Cells `x*G+y` of the reference's `generate_random_positions` (BASE:156-177): draws from
    `np.random.default_rng(seed)` go into a set until it holds n_total positions; the set's
    iteration order is the placement order (predators, prey, grass, BASE:185-187)."""
    if n_total > grid_size * grid_size:
        raise ValueError("Cannot place more unique positions than grid cells.")  # BASE:167-168
    rng = np.random.default_rng(seed)
    seen = set()
    while len(seen) < n_total:
        seen.add(tuple(rng.integers(0, grid_size, size=2)))
    return np.array([int(p[0]) * grid_size + int(p[1]) for p in seen], np.int32)


class PredPreyGrass(_Base):
    """`PredPreyGrass(config)` — see module docstring.  Extra, optional config keys (ignored by the
    reference): "reward_variant" in {"sparse","eating","dense","additive","kickback"} selecting the
    project_reward_shaping variant (default "sparse" = BASE), "cuda_device" (default 0),
    "cap_live" (device slot capacity per species)."""

    reward_variant = "sparse"

    def __init__(self, config=None):
        super().__init__()
        from .batched import BatchedPredPreyGrass  # needs the CUDA extension; fails loudly without it

        config = dict(config or {}) or dict(_default_config())
        self.config = config
        g = config.get
        self.max_steps = g("max_steps", 10000)
        self.grid_size = g("grid_size", 10)
        self.num_obs_channels = g("num_obs_channels", 4)
        self.predator_obs_range = g("predator_obs_range", 7)
        self.prey_obs_range = g("prey_obs_range", 5)
        self.n_possible_predators = g("n_possible_predators", 50)
        self.n_possible_prey = g("n_possible_prey", 50)
        self.n_initial_active_predator = g("n_initial_active_predator", 6)
        self.n_initial_active_prey = g("n_initial_active_prey", 8)
        self.initial_num_grass = g("initial_num_grass", 25)
        variant = g("reward_variant", self.reward_variant)
        if variant not in REWARD_MODES:
            raise ValueError(f"unknown reward_variant {variant!r}")
        cap = g("cap_live", None)
        self._cfg = make_config(config, reward_mode=variant, cap_live=cap, autoreset=False, seed=g("seed", 0) or 0)
        self._batch = BatchedPredPreyGrass(self._cfg, 1, device=g("cuda_device", 0))
        self.possible_agents = [f"predator_{i}" for i in range(self.n_possible_predators)] + [
            f"prey_{j}" for j in range(self.n_possible_prey)]
        self.agents = [f"predator_{i}" for i in range(self.n_initial_active_predator)] + [
            f"prey_{j}" for j in range(self.n_initial_active_prey)]
        self.grass_agents = [f"grass_{k}" for k in range(self.initial_num_grass)]
        C = self.num_obs_channels
        pred_space = Box(low=0.0, high=100.0, shape=(C, self.predator_obs_range, self.predator_obs_range), dtype=np.float64)
        prey_space = Box(low=0.0, high=100.0, shape=(C, self.prey_obs_range, self.prey_obs_range), dtype=np.float64)
        self.observation_spaces = {a: pred_space if "predator" in a else prey_space for a in self.possible_agents}
        self.action_to_move_tuple = {a: (a // 3 - 1, a % 3 - 1) for a in range(9)}  # BASE:96-106
        self.num_actions = 9
        act_space = Discrete(9)
        self.action_spaces = {a: act_space for a in self.possible_agents}
        self.agents_just_ate = set()
        self.cumulative_rewards = {}
        self.current_step = 0
        self._rows = {}            # agent id string -> (species, row) in the last output (live agents only)
        self._trunc_pending = None  # observations to hand out on BASE's extra truncation call
        self._done = True
        self._state = None

    # ------------------------------------------------------------------ reset / step
    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        b = self._batch
        n_total = self.n_initial_active_predator + self.n_initial_active_prey + self.initial_num_grass
        # options={"ppg_tape": cells}: recorded spawn-fallback cells (BASE:760-764 draws them from the GLOBAL numpy generator,
        # which no seed of this class reaches) replayed after the reset cells; without it the device's Philox stream decides
        extra = np.asarray((options or {}).get("ppg_tape", ()), np.int32).reshape(-1)
        if seed is not None:
            b.load_tape([np.concatenate([reference_initial_cells(seed, self.grid_size, n_total), extra])])
            b.reset(seeds=np.array([np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)], np.uint64))
        else:
            b.load_tape([np.zeros(0, np.int32)])
            b.reset(seeds=np.array([np.random.SeedSequence().generate_state(1, np.uint64)[0]], np.uint64))
        out = b.outputs_numpy()
        raise_on_status(int(out["env_status"][0]))
        self.current_step = 0
        self._trunc_pending = None
        self._done = False
        self._state = None
        self.agents_just_ate = set()
        obs, _, _, _, names = self._dicts(out)
        self.agents = list(names)
        self.cumulative_rewards = {a: 0 for a in self.agents}
        return obs, {}

    def step(self, action_dict):
        import torch

        if self._done and self._trunc_pending is None:
            raise RuntimeError("step() called on a finished episode; call reset()")
        self._state = None
        if self._trunc_pending is not None:
            # BASE:228-238 — the call after max_steps real steps: same state, everybody truncated
            pend = self._trunc_pending
            self._trunc_pending = None
            self._done = True
            self.agents = [a for a in self.agents if a in pend]  # BASE:222-225
            obs = {a: pend[a] for a in self.agents}              # BASE:229 iterates the sorted self.agents
            rewards = {a: 0.0 for a in obs}
            trunc = {a: True for a in obs}
            term = {a: False for a in obs}
            trunc["__all__"], term["__all__"] = True, False
            return obs, rewards, term, trunc, {}
        b = self._batch
        acts = [np.full(max(1, b.row_capacity[s]), 4, np.int32) for s in range(2)]
        order = [np.zeros(max(1, b.row_capacity[s]), np.int32) for s in range(2)]
        seen = [0, 0]
        for agent, action in action_dict.items():
            if agent not in self._rows:
                raise KeyError(agent)  # BASE:246: the reference indexes agent_energies[agent]
            s, row = self._rows[agent]
            self.action_to_move_tuple[int(action)]  # KeyError on an action outside the space, as BASE:502
            acts[s][row] = int(action)
            order[s][row] = seen[s]
            seen[s] += 1
        if seen[0] + seen[1] != len(self._rows):
            missing = [a for a in self._rows if a not in action_dict]
            raise KeyError(f"action_dict misses live agents {missing[:4]} (every live agent must act)")
        dev = b.device
        t = [torch.from_numpy(x).to(dev) for x in acts + order]
        b.step_ordered(t[0], t[1], t[2], t[3])
        out = b.outputs_numpy()
        raise_on_status(int(out["env_status"][0]))
        obs, rew, term, trunc, names = self._dicts(out)
        self.current_step = int(out["env_step"][0])
        flags = int(out["env_flags"][0])
        term["__all__"] = bool(flags & ENV_TERMINATED)
        trunc["__all__"] = False  # BASE:463 — truncation is only ever reported by the extra call
        self.agents = sorted(names)  # BASE:468 (terminated ones stay listed until the next call)
        for a, r in rew.items():
            self.cumulative_rewards[a] = self.cumulative_rewards.get(a, 0) + r
        if flags & ENV_TERMINATED:
            self._done = True
        elif flags & ENV_TRUNCATED:
            self._trunc_pending = {a: obs[a] for a in obs if a in self._rows}
            self._done = True
        return obs, rew, term, trunc, {}

    def _dicts(self, out):
        """row batch of env 0 -> the reference's dicts, in its observation-dict order
        (old predators, old prey, newborn predators, newborn prey: BASE:459 over self.agents)."""
        obs, rew, term, trunc = {}, {}, {}, {}
        self._rows = {}
        self.agents_just_ate = set()
        for group in ("old", "new"):
            for s in range(2):
                if group == "old":
                    r0, r1 = int(out[f"old_off{s}"][0]), int(out[f"old_off{s}"][1])
                else:
                    r0 = int(out[f"new_off{s}"][0])
                    r1 = r0 + int(out[f"new_cnt{s}"][0])
                for r in range(r0, r1):
                    name = f"{_SPECIES[s]}_{int(out[f'row_agent{s}'][r])}"
                    f = int(out[f"flags{s}"][r])
                    obs[name] = out[f"obs{s}"][r].astype(np.float64)
                    rew[name] = float(out[f"reward{s}"][r])
                    term[name] = bool(f & ROW_TERMINATED)
                    trunc[name] = False
                    if f & ROW_ATE:
                        self.agents_just_ate.add(name)
                    if not f & ROW_TERMINATED:
                        self._rows[name] = (s, r)
        return obs, rew, term, trunc, list(obs)

    # ------------------------------------------------------------------ attributes read by renderers
    def _read(self):
        if self._state is None:
            self._state = self._batch.read_env(0)
        return self._state

    @property
    def agent_positions(self):
        st = self._read()
        return {f"{_SPECIES[s]}_{int(i)}": (int(x), int(y)) for s in range(2) for i, (x, y) in zip(st["ids"][s], st["xy"][s])}

    @property
    def predator_positions(self):
        return {k: v for k, v in self.agent_positions.items() if k.startswith("predator")}

    @property
    def prey_positions(self):
        return {k: v for k, v in self.agent_positions.items() if k.startswith("prey")}

    @property
    def agent_energies(self):
        st = self._read()
        return {f"{_SPECIES[s]}_{int(i)}": float(e) for s in range(2) for i, e in zip(st["ids"][s], st["energy"][s])}

    @property
    def grass_positions(self):
        st = self._read()
        return {f"grass_{k}": (int(x), int(y)) for k, (x, y) in enumerate(st["grass_xy"])}

    @property
    def grass_energies(self):
        st = self._read()
        return {f"grass_{k}": float(e) for k, e in enumerate(st["grass_energy"])}

    @property
    def current_num_predators(self):
        return len(self._read()["ids"][0])

    @property
    def current_num_prey(self):
        return len(self._read()["ids"][1])

    @property
    def grid_world_state(self):
        """[C, G, G] float64 view of the world for renderers: rebuilt from the agent lists the way
        the device rebuilds it at the start of a step (channel 0 is the empty wall layer)."""
        st = self._read()
        g = np.zeros((self.num_obs_channels, self.grid_size, self.grid_size), np.float64)
        for s in range(2):
            for (x, y), e in zip(st["xy"][s], st["energy"][s]):
                g[1 + s, x, y] = e
        for (x, y), e in zip(st["grass_xy"], st["grass_energy"]):
            g[3, x, y] = e
        return g

    # ------------------------------------------------------------------ snapshot / restore (BASE:768-804)
    def get_state_snapshot(self):
        return {"blob": self._batch.snapshot(), "current_step": self.current_step, "agents": list(self.agents),
                "rows": dict(self._rows), "agents_just_ate": set(self.agents_just_ate),
                "cumulative_rewards": dict(self.cumulative_rewards), "done": self._done,
                "trunc_pending": None if self._trunc_pending is None else dict(self._trunc_pending)}

    def restore_state_snapshot(self, snapshot):
        self._batch.restore(snapshot["blob"])
        self.current_step = snapshot["current_step"]
        self.agents = list(snapshot["agents"])
        self.agents_just_ate = set(snapshot["agents_just_ate"])
        self.cumulative_rewards = dict(snapshot["cumulative_rewards"])
        self._done = snapshot["done"]
        self._trunc_pending = None if snapshot["trunc_pending"] is None else dict(snapshot["trunc_pending"])
        self._state = None
        # restore relabels the rows densely in list order (ppg_restore): re-derive the id -> row map
        out = self._batch.outputs_numpy()
        self._rows = {}
        for s in range(2):
            off = out[f"old_off{s}"]
            for r in range(int(off[0]), int(off[1])):
                self._rows[f"{_SPECIES[s]}_{int(out[f'row_agent{s}'][r])}"] = (s, r)

    def close(self):
        if getattr(self, "_batch", None) is not None:
            self._batch.close()
            self._batch = None


class PredPreyGrassDenseRewards(PredPreyGrass):
    """project_reward_shaping/base_environment_dense_rewards"""
    reward_variant = "dense"


class PredPreyGrassDenseRewardsAdditive(PredPreyGrass):
    """project_reward_shaping/base_environment_dense_rewards_additive (BASELINE configs[2])"""
    reward_variant = "additive"


class PredPreyGrassSparseRewardsPlusKickback(PredPreyGrass):
    """project_reward_shaping/base_environment_sparse_rewards_plus_kickback"""
    reward_variant = "kickback"


class PredPreyGrassSparseRewards(PredPreyGrass):
    """project_reward_shaping/base_environment_sparse_rewards: the same class as BASE (the reference copy differs in its
    docstring and config import only)"""
    reward_variant = "sparse"


class PredPreyGrassSparseRewardsPlusEating(PredPreyGrass):
    """project_reward_shaping/base_environment_sparse_rewards_plus_eating: BASE's code with the eating rewards of its own
    config_env.py (`reward_predator_catch_prey`, `reward_prey_eat_grass`: BASE:322,365 pay them when non-zero)"""
    reward_variant = "eating"


class PredPreyGrassSeasonal(PredPreyGrass):
    """non_evolutionary/base_environment_seasonal: BASE with a square-wave multiplier on the grass regrowth
    (`season_length_steps`, `season_high_multiplier`, `season_low_multiplier`; SEASON:63-67,224-234,268-271 with
    SEASON = predpreygrass/non_evolutionary/base_environment_seasonal/predpreygrass_rllib_env.py)."""
    reward_variant = "seasonal"

    def __init__(self, config=None):
        from .config import SEASONAL_CONFIG

        config = dict(config or SEASONAL_CONFIG)
        config.setdefault("season_length_steps", 40)  # the variant's own defaults (SEASON:64-66)
        super().__init__(config)
        self.season_length_steps = config["season_length_steps"]
        self.season_high_multiplier = config.get("season_high_multiplier", 1.5)
        self.season_low_multiplier = config.get("season_low_multiplier", 0.5)

    def _current_season_multiplier(self):
        phase = (self.current_step // self.season_length_steps) % 2
        return self.season_high_multiplier if phase == 0 else self.season_low_multiplier


def _default_config():
    from .config import BASE_CONFIG

    return BASE_CONFIG
