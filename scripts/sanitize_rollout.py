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
host = env.make_host_buffers(pinned=True)
n0, n1 = env.out.counts()
host["actions0"][:n0] = 4
host["actions1"][:n1] = 4
env.step_host(host)
torch.cuda.synchronize()
st = env.stats()
print(which, "ok", {k: st[k] for k in ("env_steps", "agent_steps", "episodes", "births_pred", "births_prey", "status_envs")})
env.close()
