#!/usr/bin/env python
"""Record `per_step_agent_data` (MR:391-411) and `agent_event_log` (MR:1153-1165, 775-783, 825-833, 877-880, 917-925) of the
UNMODIFIED reference classes of the trait variants (metabolic_rate / investment / cooperation) for some of the existing
mr_* / inv_* / coop_* recordings, plus the order in which the episode's agent records are iterated
(`_iter_all_agent_records`, MR:1400-1404: live records, then completed ones in the order they were finalized) — the
`*_repro_spearman` metrics depend on it (MR:1358-1382).

Runs in the build container only (needs /root/reference).  Same method as make_golden_eco_events.py: the recorded action
dicts are replayed through the reference class (tests/golden/_shim stubs), the observation hashes are checked against the
recording on the way, and the exporters are written to tests/golden/trait_events_<case>.json.gz.
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

PKG = {"mr": "predpreygrass.evolutionary.eco_evolutionary_metabolic_rate",
       "inv": "predpreygrass.evolutionary.eco_evolutionary_investment",
       "coop": "predpreygrass.evolutionary.eco_evolutionary_cooperation",
       "cad": "predpreygrass.evolutionary.eco_evolutionary_cadence"}
NAMES = ("predator", "prey")
MAX_STEPS = {"mr_density_s5": 80, "mr_crowded_s1": 50, "inv_crowded_s4": 58, "coop_crowded_s3": 86, "mr_rich_s3": 60, "coop_share_s2": 40, "inv_nogenome_s6": 40, "cad_rich_s3": 60, "cad_crowded_s4": 60, "cad_nogenome_s6": 50}
CASES = ("mr_default_s1", "mr_crowded_s1", "mr_density_s5", "mr_nogenome_s1", "mr_trunc_s2", "mr_rich_s3",
         "inv_default_s1", "inv_crowded_s4", "inv_nogenome_s6",
         "coop_default_s1", "coop_crowded_s3", "coop_trunc_s4", "coop_share_s2",
         "cad_default_s1", "cad_aging_s5", "cad_crowded_s4", "cad_rich_s3", "cad_nogenome_s6", "cad_trunc_s7")


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
    for case in (sys.argv[1:] or CASES):
        fam = case.split("_")[0]
        mod = importlib.import_module(PKG[fam] + ".predpreygrass_rllib_env")
        z = np.load(os.path.join(HERE, case + ".npz"))
        cfg = json.loads(str(z["cfg_json"]))
        cfg.pop("variant", None)
        if fam == "cad":
            cfg["record_step_data"] = True  # CAD:83,422: per_step_agent_data is optional there
        env = mod.PredPreyGrass(cfg)
        env.reset(seed=int(z["seed"]))
        n, ended, metrics = 0, False, None
        for t in range(min(len(z["steps"]), MAX_STEPS.get(case, 10 ** 9))):
            a0, a1 = z["act_off"][t], z["act_off"][t + 1]
            acts = {f"{NAMES[s]}_{i}": int(v) for s, i, v in zip(z["act_s"][a0:a1], z["act_id"][a0:a1], z["act_v"][a0:a1])}
            obs, rew, term, trunc, infos = env.step(acts)
            keys = sorted(obs, key=lambda a: (a.startswith("prey"), int(a.rsplit("_", 1)[1])))
            win = (lambda o: o["observations"]) if fam == "cad" else (lambda o: o)
            assert np.array_equal(sha([win(obs[k]) for k in keys]), z["obs_sha"][t]), (case, t)  # same episode as the recording
            n = t + 1
            if term["__all__"] or trunc["__all__"]:
                ended = True
                metrics = infos["__all__"]["training_metrics"]
                break
        out = {"steps": n, "ended": ended, "per_step_agent_data": plain(env.per_step_agent_data), "agent_event_log": plain(env.agent_event_log),
               "record_order": [a for a, _ in env._iter_all_agent_records()],
               "agent_stats": plain(env.get_all_agent_stats()), "energy_by_type": plain(env.get_total_energy_by_type()),
               "offspring_by_type": plain(env.get_total_offspring_by_type()),
               "spearman": {k: float(v) for k, v in (metrics or {}).items() if "spearman" in k}}
        path = os.path.join(HERE, f"trait_events_{case}.json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(out, sort_keys=True).encode())
        ev = env.agent_event_log
        print(case, "steps", n, "ended", ended, "agents", len(ev), "eating", sum(len(e["eating_events"]) for e in ev.values()),
              "repro", sum(len(e["reproduction_events"]) for e in ev.values()), out["spearman"], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
