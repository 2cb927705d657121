"""SASS instruction histogram per kernel of the shipped library (no GPU needed): the mnemonics that show which hardware paths
a kernel uses — UBLKCP (bulk-copy engine), SYNCS (mbarrier), STG.E.EF.128 (evict-first 16-byte stores), ATOM / RED,
MEMBAR, LDS / STS, DFMA-class fp64.  usage: python scripts/sass_histogram.py > profiles/r02_sass_histogram.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "predpreygrass_b200", "libppg_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
kern, cnt = None, {}
for l in txt:
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(ppg::StepParams\)|void |ppg::", "", kern)
        cnt[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and kern:
        cnt[kern][m.group(1)] += 1
KEYS = [("UBLKCP", "bulk copy"), ("SYNCS", "mbarrier"), ("STG.E.EF.128", "evict-first v4 stores"), ("STG", "global stores"), ("LDG", "global loads"),
        ("ATOM", "atomics (returning)"), ("RED", "reductions"), ("MEMBAR", "fences"), ("LDS", "shared loads"), ("STS", "shared stores"),
        ("WARPSYNC", "warp syncs"), ("SHFL", "shuffles"), ("DADD", "fp64 add"), ("DMUL", "fp64 mul"), ("DFMA", "fp64 fma"), ("BAR", "CTA barriers")]
print("# SASS instruction histogram of predpreygrass_b200/libppg_b200.so (sm_100a), per kernel\n")
print("| kernel | instructions | " + " | ".join(k for k, _ in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for k, c in sorted(cnt.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(c.values())
    if tot < 40:
        continue
    row = []
    for key, _ in KEYS:
        row.append(sum(v for m, v in c.items() if m == key or m.startswith(key + ".") or (key == "STG.E.EF.128" and m.startswith("STG.E.EF.128"))))
    print(f"| `{k[:110]}` | {tot} | " + " | ".join(str(v) for v in row) + " |")
print("\n" + "; ".join(f"{k} = {d}" for k, d in KEYS))
