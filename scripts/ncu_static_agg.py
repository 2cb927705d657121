"""static code footprint (SASS instructions) and stall samples per source line / file of one kernel
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python ncu_static_agg.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = None; stat = {}; cur = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); continue
    if len(r) < 10: continue
    if r[0] != '':
        try: cur = (f, int(r[0]))
        except ValueError: continue
        stat.setdefault(cur, [r[1], 0, 0, 0])
    if r[2].startswith('0x') and cur is not None:
        a = stat[cur]; a[1] += 1
        try: a[2] += float(r[iS] or 0); a[3] += float(r[iI] or 0)
        except ValueError: pass
tot = sum(a[1] for a in stat.values())
print("static SASS instructions", tot, "=", tot * 16 / 1024, "KB")
byfile = {}
for (f, l), a in stat.items():
    b = byfile.setdefault(f, [0, 0, 0]); b[0] += a[1]; b[1] += a[2]; b[2] += a[3]
for f, b in sorted(byfile.items(), key=lambda x: -x[1][0]): print(f"{f:32s} static {b[0]:6d}  samples {b[1]:7.0f}  executed {b[2]:10.0f}")
print("--- top lines by static count")
for (f, l), a in sorted(stat.items(), key=lambda x: -x[1][1])[:N]:
    print(f"{f}:{l} static={a[1]} samples={a[2]:.0f} dyn={a[3]:.0f} | {a[0][:90]}")
