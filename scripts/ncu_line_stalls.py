"""Per source line: warp-stall samples with the dominant stall reasons, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`.  usage: ncu_line_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg, f, hdr = {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        st = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if r[0] == "" or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0]); s = float(r[iS] or 0); i = float(r[iI] or 0)
    except ValueError:
        continue
    a = agg.setdefault((f, ln), [r[1], 0.0, 0.0, {}])
    a[1] += s; a[2] += i
    for k, h in st:
        try:
            v = float(r[k] or 0)
        except ValueError:
            v = 0
        if v:
            a[3][h] = a[3].get(h, 0) + v
ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
print("total samples", ts, "warp instructions", ti)
tot_st = {}
for a in agg.values():
    for h, v in a[3].items():
        tot_st[h] = tot_st.get(h, 0) + v
print("stall totals:", ", ".join(f"{h[6:]} {100 * v / ts:.1f}%" for h, v in sorted(tot_st.items(), key=lambda x: -x[1])[:10]))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:N]:
    top = ", ".join(f"{h[6:]} {v:.0f}" for h, v in sorted(a[3].items(), key=lambda x: -x[1])[:3])
    print(f"{k[0][:22]}:{k[1]:4d} s={a[1]:6.0f} ({100 * a[1] / ts:4.1f}%) i={a[2]:8.0f} [{top}] {a[0].strip()[:90]}")
