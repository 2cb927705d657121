"""Replay an ECO golden case on the GPU and on the oracle side by side; print the first difference (debug aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle.oracle import Oracle
from predpreygrass_b200.batched import BatchedPredPreyGrass
from tests.helpers import config_from_golden, load_golden

name = sys.argv[1]
z, cfg = load_golden(name)
c = config_from_golden(cfg, autoreset=False)
c.cap_live[0] = min(c.cap_live[0], 224); c.cap_live[1] = min(c.cap_live[1], 416)
g = BatchedPredPreyGrass(c, 1)
o = Oracle(c, 1)
g.load_tape([np.concatenate([z["init_cells"], z["fallback_cells"]])], [np.concatenate([z["founder_speed"], z["step_reals"]])])
o.load_tape([z["fallback_cells"]], [z["step_reals"]])
g.reset()
oo = o.env_reset_eco(0, z["init_cells"], z["founder_speed"])
out = g.outputs_numpy()
for t in range(len(z["steps"])):
    a0, a1 = z["act_off"][t], z["act_off"][t + 1]
    act, rank, seen = {}, {}, [0, 0]
    for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1]):
        act[(int(s), int(i))] = int(v); rank[(int(s), int(i))] = seen[int(s)]; seen[int(s)] += 1
    orders = []
    for s in range(2):
        n = out["n"][s]
        a = np.full(max(n, 1), 12, np.int32); od = np.zeros(max(n, 1), np.int32)
        for r in range(n):
            if not (out[f"flags{s}"][r] & 1):
                a[r] = act[(s, int(out[f"row_agent{s}"][r]))]; od[r] = rank[(s, int(out[f"row_agent{s}"][r]))]
        g.actions[s][: len(a)].copy_(torch.from_numpy(a)); orders.append(torch.from_numpy(od).cuda())
    g.step_ordered(g.actions[0], g.actions[1], orders[0], orders[1])
    oo = o.env_step_ordered(0, z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])
    out = g.outputs_numpy()
    bad = False
    for s in range(2):
        for k in ("row_agent", "flags", "reward"):
            if not np.array_equal(out[f"{k}{s}"], oo[f"{k}{s}"]):
                print("step", t, k, s, "differs", out[f"{k}{s}"], oo[f"{k}{s}"]); bad = True
        if out[f"obs{s}"].shape == oo[f"obs{s}"].shape and not np.array_equal(out[f"obs{s}"], oo[f"obs{s}"]):
            d = np.argwhere(out[f"obs{s}"] != oo[f"obs{s}"])
            print("step", t, "obs", s, "differs at", d[:12].tolist(), "agents", out[f"row_agent{s}"][np.unique(d[:, 0])][:10], "flags", out[f"flags{s}"][np.unique(d[:, 0])][:10])
            print(" gpu", out[f"obs{s}"][tuple(d[:8].T)], "ora", oo[f"obs{s}"][tuple(d[:8].T)]); bad = True
    if out["env_flags"][0] & 3:
        break
    a, b = g.read_env_eco(0), o.read_env_eco(0)
    for s in range(2):
        for k in ("ids", "xy", "energy", "age", "speed"):
            if not np.array_equal(a[k][s], b[k][s]):
                print("step", t, "state", k, s, "differs"); print(a[k][s]); print(b[k][s]); bad = True
    if not np.array_equal(a["dead_prey"], b["dead_prey"]): print("dead differs", a["dead_prey"], b["dead_prey"]); bad = True
    if not np.array_equal(a["grass_energy"], b["grass_energy"]): print("grass differs"); bad = True
    if bad:
        break
print("done at step", t, "status", out["env_status"], oo["env_status"])
