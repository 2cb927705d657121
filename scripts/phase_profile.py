"""Per-phase latency of the step kernel, mean SM cycles per env-step (experimental build, scripts/build_prof.py):
PPG_LIB=predpreygrass_b200/libppg_b200_prof.so python scripts/phase_profile.py --variant base --envs 4096"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench

args = bench.parse()
import torch

from predpreygrass_b200 import _lib
from predpreygrass_b200.batched import BatchedPredPreyGrass

NAMES = {
    "base": ["ticket", "hdr+prefix", "reset", "load lists", "load grass", "owner maps", "movement", "sort", "predators", "prey", "births",
             "counts", "publish", "refresh tables", "rows pass", "dump image", "unwrite+grass store", "hdr store+counters"],
}
fam = "stag" if args.variant == "stag" else ("eco" if args.variant in bench.ECO_FAMILY else "base")
L = _lib.load()
fn = getattr(L, f"ppg_debug_phase_cycles_{fam}")
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
cfg = bench.build_config(args, seed=1000)
env = BatchedPredPreyGrass(cfg, args.envs, device=0)
env.reset()
for _ in range(args.preroll):
    a0, a1 = env.random_actions(4242); env.step(a0, a1)
torch.cuda.synchronize()
K = 20
acc = []
for _ in range(K):
    a0, a1 = env.random_actions(4242); env.step(a0, a1)
    torch.cuda.synchronize()
    buf = np.zeros((args.envs, 24), np.uint32)
    assert fn(buf.ctypes.data, args.envs) == 0
    acc.append(buf.astype(np.float64))
acc = np.concatenate(acc)          # [K * envs][24] cycles per phase of one env-step
tot_env = acc.sum(1)
names = NAMES.get(fam, [])


def show(sel, title):
    per = acc[sel].mean(0)
    tot = per.sum()
    print(f"{title}: {int(sel.sum())} env-steps, {tot:.0f} cycles per env-step ({tot / 1.965e3:.1f} us at 1965 MHz)")
    for k, v in enumerate(per):
        if v > 0:
            print(f"  {k:2d} {names[k] if k < len(names) else '':24s} {v:8.0f}  {100 * v / tot:5.1f} %")


print(f"{args.variant} {args.envs} envs; percentiles of cycles per env-step [10, 50, 90, 99, 99.9, 100]:",
      np.percentile(tot_env, [10, 50, 90, 99, 99.9, 100]).astype(int).tolist())
show(np.ones(len(acc), bool), "all")
show(tot_env >= np.percentile(tot_env, 99), "slowest 1 %")
