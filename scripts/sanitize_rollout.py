"""Small rollout of every variant for compute-sanitizer (memcheck / racecheck / synccheck): 256 envs, resets, births,
deaths, newborn rows, auto-reset, snapshot / restore and the host-buffer step — every kernel of the library runs."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from predpreygrass_b200.batched import BatchedPredPreyGrass  # noqa: E402
from predpreygrass_b200.config import BASE_CONFIG, ECO_CONFIG, STAG_CONFIG, VARIANT_ECO, VARIANT_STAG, make_config  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "base"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
envs = int(sys.argv[3]) if len(sys.argv) > 3 else 256
if which == "base":
    cfg = make_config(dict(BASE_CONFIG, max_steps=25), reward_mode="kickback", cap_live=(32, 128), seed=3)
elif which == "add":
    cfg = make_config(dict(BASE_CONFIG, max_steps=25), reward_mode="additive", cap_live=(32, 128), seed=3)
elif which == "eco":
    cfg = make_config(dict(ECO_CONFIG, max_steps=25, energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0,
                           predator_creation_energy_threshold=8.0, max_energy_gain_per_prey=2.0), variant=VARIANT_ECO, cap_live=(64, 128), seed=3)
elif which == "eco_lean":  # the shipped config: the step kernel without the carcass / ghost-cell / juvenile code (ppg_eco.cu KIND 3)
    cfg = make_config(dict(ECO_CONFIG, max_steps=25), variant=VARIANT_ECO, cap_live=(64, 128), seed=3)
elif which.endswith("_ev"):  # trait variants with the episode's event counters on (ppg_read_episode_events_eco): tiny id pools, a
    # tight density cap and a satiation cooldown so that every counter branch of the kernel runs
    from predpreygrass_b200.config import TRAIT_CONFIGS  # noqa: E402

    base = which[:-3]
    extra = {"metabolic": dict(predator_reproduction_max_ratio=0.3, predator_satiation_cooldown=6), "investment": dict(predator_satiation_cooldown=6),
             "cooperation": dict(cooperation_range=3)}[base]
    cfg = make_config(dict(TRAIT_CONFIGS[base], max_steps=25, energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0,
                           predator_creation_energy_threshold=8.0, n_possible_predators=24, n_possible_prey=40,
                           genome_mutation={"rate": 1.0, "std": 0.3}, **extra), variant=VARIANT_ECO, cap_live=(32, 64), seed=3,
                      track_episode_sums=True)
elif which in ("metabolic", "investment", "cooperation", "cadence"):
    from predpreygrass_b200.config import TRAIT_CONFIGS  # noqa: E402

    cfg = make_config(dict(TRAIT_CONFIGS[which], max_steps=25, energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0,
                           predator_creation_energy_threshold=8.0), variant=VARIANT_ECO, cap_live=(64, 160), seed=3)
else:
    cfg = make_config(dict(STAG_CONFIG, max_steps=25), variant=VARIANT_STAG, cap_live=(64, 192), seed=3)
env = BatchedPredPreyGrass(cfg, envs, device=0)
env.reset()
for t in range(steps):
    a0, a1 = env.random_actions(99)
    env.step(a0, a1)
    if t == steps // 2:
        blob = env.snapshot()
        env.restore(blob)
    if t == 23 and which.endswith("_ev"):  # the counters of the first episode, one step before its time limit
        ev = [env.read_episode_events_eco(e) for e in range(envs)]
        first_episode = {"blocked_capacity": sum(sum(v["blocked_capacity"]) for v in ev), "blocked_density": sum(v["blocked_density"] for v in ev),
                         "satiation_blocked": sum(v["satiation_blocked"] for v in ev), "donated": round(sum(sum(v["donated"]) for v in ev), 3)}
host = env.make_host_buffers(pinned=True)
n0, n1 = env.out.counts()
host["actions0"][:n0] = 4
host["actions1"][:n1] = 4
env.step_host(host)
torch.cuda.synchronize()
st = env.stats()
extra_out = first_episode if which.endswith("_ev") else {}
print(which, "ok", {k: st[k] for k in ("env_steps", "agent_steps", "episodes", "births_pred", "births_prey", "status_envs")}, extra_out)
env.close()
