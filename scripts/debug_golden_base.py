"""replay one BASE-family golden trajectory on the GPU and print the first state mismatch in detail"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from predpreygrass_b200.batched import BatchedPredPreyGrass
from tests.helpers import config_from_golden, load_golden

name = sys.argv[1]
z, cfg = load_golden(name)
c = config_from_golden(cfg, autoreset=False)
c.cap_live[0] = min(c.cap_live[0], 320); c.cap_live[1] = min(c.cap_live[1], 320)
g = BatchedPredPreyGrass(c, 1)
g.load_tape([np.concatenate([z["init_cells"], z["fallback_cells"]])])
g.reset()
out = g.outputs_numpy()
prev = g.read_env(0)
for t in range(len(z["steps"])):
    a0, a1 = z["act_off"][t], z["act_off"][t + 1]
    act, rank, seen = {}, {}, [0, 0]
    for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1]):
        act[(int(s), int(i))] = int(v); rank[(int(s), int(i))] = seen[int(s)]; seen[int(s)] += 1
    orders = []
    for s in range(2):
        n = out["n"][s]
        a = np.full(max(n, 1), 4, np.int32); o = np.zeros(max(n, 1), np.int32)
        for r in range(n):
            if not (out[f"flags{s}"][r] & 1):
                a[r] = act[(s, int(out[f"row_agent{s}"][r]))]; o[r] = rank[(s, int(out[f"row_agent{s}"][r]))]
        g.actions[s][: len(a)].copy_(torch.from_numpy(a)); orders.append(torch.from_numpy(o).cuda())
    if str(z["order"]) == "shuffle":
        g.step_ordered(g.actions[0], g.actions[1], orders[0], orders[1])
    else:
        g.step()
    out = g.outputs_numpy()
    if out["env_flags"][0] & 3:
        break
    st = g.read_env(0)
    s0, s1 = z["st_off"][t], z["st_off"][t + 1]
    bad = False
    for s in range(2):
        m = z["st_s"][s0:s1] == s
        ref = list(zip(z["st_id"][s0:s1][m].tolist(), z["st_x"][s0:s1][m].tolist(), z["st_y"][s0:s1][m].tolist(), z["st_e"][s0:s1][m].tolist()))
        mine = list(zip(st["ids"][s].tolist(), st["xy"][s][:, 0].tolist(), st["xy"][s][:, 1].tolist(), st["energy"][s].tolist()))
        if ref != mine:
            bad = True
            print("step", t, "species", s)
            print(" before:", list(zip(prev["ids"][s].tolist(), prev["xy"][s][:, 0].tolist(), prev["xy"][s][:, 1].tolist(), [round(e, 3) for e in prev["energy"][s].tolist()])))
            print(" other species before:", list(zip(prev["ids"][1 - s].tolist(), prev["xy"][1 - s][:, 0].tolist(), prev["xy"][1 - s][:, 1].tolist())))
            print(" actions in dict order:", [(int(ss), int(i), int(v)) for ss, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1]) if ss == s])
            print(" ref :", [(i, x, y, round(e, 3)) for i, x, y, e in ref])
            print(" mine:", [(i, x, y, round(e, 3)) for i, x, y, e in mine])
            print(" status", out["env_status"][0], "G", c.grid_size)
    if bad:
        break
    prev = st
else:
    print("no mismatch")
g.close()
