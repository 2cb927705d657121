#!/usr/bin/env python
"""Initial placements of the UNMODIFIED reference `reset(seed)` for seeds 0..N-1 (BASE default config).

These are the reset part of the replay tape for the 4096-env parity run (BASELINE configs[1]):
positions come out of numpy's PCG64 + CPython set ordering (BASE:156-187) and are recorded, not
re-derived.  Writes tests/golden/base_reset_cells_4096.npz (cells = x*grid_size+y, order
predators, prey, grass).  Build container only (needs /root/reference).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, os.environ.get("PPG_REFERENCE", "/root/reference"))

from predpreygrass.non_evolutionary.base_environment.predpreygrass_rllib_env import PredPreyGrass  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = PredPreyGrass()
G = env.grid_size
cells = np.zeros((N, env.n_initial_active_predator + env.n_initial_active_prey + env.initial_num_grass), np.int16)
for seed in range(N):
    env.reset(seed=seed)
    c = [int(env.agent_positions[a][0]) * G + int(env.agent_positions[a][1]) for a in env.agents]
    c += [int(p[0]) * G + int(p[1]) for p in env.grass_positions.values()]
    cells[seed] = c
np.savez_compressed(os.path.join(HERE, f"base_reset_cells_{N}.npz"), cells=cells, grid_size=np.int32(G))
print(cells.shape, cells[:2, :10])
