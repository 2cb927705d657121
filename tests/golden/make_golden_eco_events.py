#!/usr/bin/env python
"""Record `per_step_agent_data` (ECO:426-446) and `agent_event_log` (ECO:1488-1500, 866-877, 926-937, 1051-1059, 1083-1089,
1130-1134, 1172-1180) of the UNMODIFIED reference ECO class for some of the existing eco_*.npz recordings.

Runs in the build container only (needs /root/reference).  Each recording holds the seed, the config and the action dict
of every step; the episode is replayed through the reference class (tests/golden/_shim stubs) — the observation hashes are
checked against the recording on the way — and both exporters are written to tests/golden/eco_events_<case>.json.gz
(Python floats survive JSON exactly).  tests/test_gpu_dict_adapters.py compares `PredPreyGrassEco` against them.
"""
import gzip
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PPG_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REF)

ECO = "predpreygrass.evolutionary.eco_evolutionary"
NAMES = ("predator", "prey")
MAX_STEPS = {"eco_lineage_s1": 60, "eco_nogenome_s1": 50, "eco_jumps_s9": 60}  # populous episodes: the first steps only (fixture size)
CASES = ("eco_default_s1", "eco_default_s5_shuffle", "eco_carcass_s1", "eco_agecap_s2", "eco_juvenile_s3", "eco_crowded_s1",
         "eco_lineage_s1", "eco_ghost_s2", "eco_trunc_s2", "eco_nogenome_s1", "eco_jumps_s9")


def sha(arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def plain(obj):
    if isinstance(obj, dict):
        return {str(k): plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [plain(v) for v in obj]
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    return obj


def main():
    mod = importlib.import_module(ECO + ".predpreygrass_rllib_env")
    for case in CASES:
        z = np.load(os.path.join(HERE, case + ".npz"))
        cfg = json.loads(str(z["cfg_json"]))
        cfg.pop("variant", None)
        env = mod.PredPreyGrass(cfg)
        env.reset(seed=int(z["seed"]))
        n = 0
        for t in range(min(len(z["steps"]), MAX_STEPS.get(case, 10 ** 9))):
            a0, a1 = z["act_off"][t], z["act_off"][t + 1]
            acts = {f"{NAMES[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
            obs, rew, term, trunc, infos = env.step(acts)
            keys = sorted(obs, key=lambda a: (a.startswith("prey"), int(a.rsplit("_", 1)[1])))
            assert np.array_equal(sha([obs[k] for k in keys]), z["obs_sha"][t]), (case, t)  # same episode as the recording
            n = t + 1
            if term["__all__"] or trunc["__all__"]:
                break
        out = {"steps": n, "per_step_agent_data": plain(env.per_step_agent_data), "agent_event_log": plain(env.agent_event_log),
               "agent_stats": plain(env.get_all_agent_stats()), "energy_by_type": plain(env.get_total_energy_by_type()),
               "offspring_by_type": plain(env.get_total_offspring_by_type())}
        path = os.path.join(HERE, f"eco_events_{case[4:]}.json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(out, sort_keys=True).encode())
        ev = env.agent_event_log
        print(case, "steps", n, "agents", len(ev), "eating", sum(len(e["eating_events"]) for e in ev.values()),
              "repro", sum(len(e["reproduction_events"]) for e in ev.values()), "diet", sum(len(e["diet_events"]) for e in ev.values()),
              "lifecycle", sum(len(e["lifecycle_events"]) for e in ev.values()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
