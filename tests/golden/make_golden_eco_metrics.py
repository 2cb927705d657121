#!/usr/bin/env python
"""Record `infos["__all__"]["training_metrics"]` (`_build_episode_training_metrics`, ECO:1613-1661) of the UNMODIFIED
reference ECO class for the episodes of the existing eco_*.npz recordings.

Runs in the build container only (needs /root/reference).  Each recording holds the seed, the config and the action
dict of every step; the episode is replayed through the reference class (tests/golden/_shim stubs) — the observation
hashes are checked against the recording on the way — and the metrics dict of the last step is written to
tests/golden/eco_training_metrics.json.  tests/test_gpu_dict_adapters.py compares `PredPreyGrassEco` against it.
"""
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


def sha(arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def main():
    mod = importlib.import_module(ECO + ".predpreygrass_rllib_env")
    out = {}
    for fn in sorted(os.listdir(HERE)):
        if not (fn.startswith("eco_") and fn.endswith(".npz")):
            continue
        z = np.load(os.path.join(HERE, fn))
        cfg = json.loads(str(z["cfg_json"]))
        cfg.pop("variant", None)
        for k in ("max_agent_age", "carcass_only_predator_age"):
            if isinstance(cfg.get(k), dict):
                cfg[k] = {a: v for a, v in cfg[k].items()}
        env = mod.PredPreyGrass(cfg)
        env.reset(seed=int(z["seed"]))
        metrics = None
        for t in range(len(z["steps"])):
            a0, a1 = z["act_off"][t], z["act_off"][t + 1]
            acts = {f"{NAMES[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
            obs, rew, term, trunc, infos = env.step(acts)
            keys = sorted(obs, key=lambda a: (a.startswith("prey"), int(a.rsplit("_", 1)[1])))
            assert np.array_equal(sha([obs[k] for k in keys]), z["obs_sha"][t]), (fn, t)  # same episode as the recording
            if term["__all__"] or trunc["__all__"]:
                metrics = infos["__all__"]["training_metrics"]
                break
        if metrics is not None:
            out[fn[:-4]] = {k: float(v) for k, v in metrics.items()}
            print(fn, "episode ended at step", t + 1, "agents", metrics["predator_agent_count"], metrics["prey_agent_count"])
        else:
            print(fn, "recording stops before the episode ends: skipped")
    with open(os.path.join(HERE, "eco_training_metrics.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
