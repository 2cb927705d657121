"""Shared helpers of the parity tests."""
import glob
import hashlib
import json
import os

import numpy as np

from predpreygrass_b200.config import REWARD_MODES, make_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefixes=("base", "eating", "dense", "additive", "kickback", "seasonal")):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(p)[:-4]
        if name.split("_")[0] in prefixes and "reset_cells" not in name:
            out.append(name)
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = json.loads(str(z["cfg_json"]))
    return z, cfg


def config_from_golden(cfg, **kw):
    variant = cfg.get("variant", "sparse")
    if variant == "eco":
        from predpreygrass_b200.config import VARIANT_ECO

        kw.setdefault("cap_live", (cfg["n_possible_predators"], cfg["n_possible_prey"]))
        return make_config(cfg, variant=VARIANT_ECO, **kw)
    if variant in ("mr", "inv", "coop", "cad"):
        from predpreygrass_b200.config import VARIANT_ECO

        trait = {"mr": "metabolic_rate", "inv": "offspring_investment_fraction", "coop": "cooperation_rate", "cad": "cadence"}[variant]
        kw.setdefault("cap_live", (min(cfg["n_possible_predators"], 224), min(cfg["n_possible_prey"], 416)))
        return make_config(cfg, variant=VARIANT_ECO, trait=trait, **kw)
    if variant == "stag":
        from predpreygrass_b200.config import VARIANT_STAG

        kw.setdefault("cap_live", (cfg.get("n_possible_type_1_predators", 0) + cfg.get("n_possible_type_2_predators", 0),
                                   cfg.get("n_possible_type_1_prey", 0) + cfg.get("n_possible_type_2_prey", 0)))
        return make_config(cfg, variant=VARIANT_STAG, **kw)
    kw.setdefault("cap_live", (cfg.get("n_possible_predators", 50), cfg.get("n_possible_prey", 50)))
    return make_config(cfg, reward_mode=REWARD_MODES[variant], **kw)


def sha_f32(arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def id_order_rows(out, env=0):
    """Rows of one env sorted by (species, agent id): the order the ECO golden files use (the reference builds its
    dicts from Python sets, ECO:413,424, so it has no order of its own)."""
    rows = dict_order_rows(out, env)
    return sorted(rows, key=lambda sr: (sr[0], int(out[f"row_agent{sr[0]}"][sr[1]])))


def sha_f64(arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def dict_order_rows(out, env=0):
    """Rows of one env in the reference's observation-dict order:
    old predators, old prey, newborn predators, newborn prey (BASE:459 over self.agents)."""
    rows = []
    for s in range(2):
        off = out[f"old_off{s}"]
        rows += [(s, r) for r in range(int(off[env]), int(off[env + 1]))]
    for s in range(2):
        r0 = int(out[f"new_off{s}"][env])
        rows += [(s, r) for r in range(r0, r0 + int(out[f"new_cnt{s}"][env]))]
    return rows
