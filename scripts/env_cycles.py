"""per-env step latency distribution (ppg_profile_env_cycles): python scripts/env_cycles.py --variant stag --envs 8192"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
sys.argv = [sys.argv[0]] + sys.argv[1:]
args = bench.parse()
import torch
from predpreygrass_b200.batched import BatchedPredPreyGrass
cfg = bench.build_config(args, seed=1000)
env = BatchedPredPreyGrass(cfg, args.envs, device=0)
env.reset()
for _ in range(args.warmup):
    a0, a1 = env.random_actions(4242); env.step(a0, a1)
for rep in range(3):
    a0, a1 = env.random_actions(4242); env.step(a0, a1)
    cyc, info, t0, sm = env.profile_env_cycles()
    mode, births, agents = info & 0xFF, (info >> 8) & 0xFF, info >> 16
    q = np.percentile(cyc, [0, 10, 50, 90, 99, 99.9, 100])
    print("cycles pct[0,10,50,90,99,99.9,100] =", q.astype(int).tolist(), "sum/1e6 = %.1f" % (cyc.sum() / 1e6))
    for m in (1, 2):
        sel = mode == m
        if sel.any(): print("  mode", m, "n", int(sel.sum()), "mean", int(cyc[sel].mean()), "max", int(cyc[sel].max()))
    for b in range(0, 6):
        sel = (mode == 2) & (births == b)
        if sel.any(): print("  births", b, "n", int(sel.sum()), "mean cyc", int(cyc[sel].mean()), "mean agents %.1f" % agents[sel].mean())
    sel = (mode == 2) & (births >= 6)
    if sel.any(): print("  births>=6 n", int(sel.sum()), "mean cyc", int(cyc[sel].mean()))
    top = np.argsort(cyc)[-5:]
    print("  slowest:", [(int(e), int(cyc[e]), int(mode[e]), int(births[e]), int(agents[e])) for e in top])
    t0 = (t0 - t0.min()).astype(np.int64)  # ns since the first env was taken
    dur_ns = cyc / 1.965  # SM clock 1965 MHz under load
    end = t0 + dur_ns
    print("  timeline: last env taken at %.1f us, kernel busy until %.1f us; envs taken by 5 us slices:" % (t0.max() / 1e3, end.max() / 1e3),
          np.histogram(t0 / 1e3, bins=np.arange(0, end.max() / 1e3 + 5, 5))[0].tolist())
    busy = np.zeros(int(end.max() / 1e3) + 2)
    for a, b in zip(t0 / 1e3, end / 1e3):
        busy[int(a):int(b) + 1] += 1
    print("  warps busy per us (every 4th us):", busy[::4].astype(int).tolist())
    c = np.corrcoef(agents[mode == 2], cyc[mode == 2])[0, 1]
    print("  corr(agents, cycles) = %.2f" % c)
    last = np.argsort(end)[-12:]
    print("  last finishers (env, start us, dur us, agents, births, mode):", [(int(e), round(float(t0[e]) / 1e3, 1), round(float(dur_ns[e]) / 1e3, 1), int(agents[e]), int(births[e]), int(mode[e])) for e in last])
    firsts = t0 < 2000
    print("  first-wave envs: %d, mean agents %.1f; later envs: %d, mean agents %.1f" % (firsts.sum(), agents[firsts].mean(), (~firsts).sum(), agents[~firsts].mean() if (~firsts).any() else 0))
