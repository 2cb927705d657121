"""which PPG_STATUS_* bits the bench rollouts of a variant raise, and how close the populations get to cap_live"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bench import build_config
from predpreygrass_b200.batched import BatchedPredPreyGrass

class A: pass
a = A(); a.variant = sys.argv[1]; a.envs = int(sys.argv[2]); steps = int(sys.argv[3]); a.eco_rich = "--rich" in sys.argv; a.reward_mode = "sparse"
a.cap = eval(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4][0] == "[" else {"base": [64, 192], "eco": [128, 320] if a.eco_rich else [32, 96], "stag": [64, 192]}[a.variant]
cfg = build_config(a, seed=1000)
env = BatchedPredPreyGrass(cfg, a.envs)
env.reset()
mx = torch.zeros(2, dtype=torch.int32, device="cuda")
bits = torch.zeros(a.envs, dtype=torch.uint8, device="cuda")
for t in range(steps):
    a0, a1 = env.random_actions(4242)
    env.step(a0, a1)
    mx = torch.maximum(mx, env.out.env_count.max(dim=0).values)
    bits |= env.out.env_status
b = bits.cpu().numpy()
print(a.variant, "cap", a.cap, "max live", mx.tolist(), "envs with status", int((b != 0).sum()), {hex(1 << k): int(((b >> k) & 1).sum()) for k in range(8) if ((b >> k) & 1).any()})
print(env.stats())
