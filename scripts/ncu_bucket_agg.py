import csv, sys
# usage: ncu_bucket_agg.py src.csv [bucket=20] [top=30]  -- samples / instructions per file and bucket of source lines
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
agg = {}; f = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; n = len(hdr); iS = hdr.index('# Samples') - n; iI = hdr.index('Instructions Executed') - n; continue
    if len(r) < 10 or r[0] == '': continue
    try: ln = int(r[0]); s = float(r[iS] or 0); i = float(r[iI] or 0)
    except ValueError: continue
    a = agg.setdefault((f, ln // B * B), [0, 0]); a[0] += s; a[1] += i
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print('total samples', ts, 'instr', ti)
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    print(f"{k[0]}:{k[1]:4d}+ s={a[0]:7.0f} ({100*a[0]/ts:4.1f}%) i={a[1]:10.0f} ({100*a[1]/ti:4.1f}%)")
