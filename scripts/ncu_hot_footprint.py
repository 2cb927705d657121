"""Static footprint of the HOT code per source line (instructions executed by at least FRAC of the envs), from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`.  usage: ncu_hot_footprint.py src.csv B [FRAC] [N]"""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1])); B = int(sys.argv[2]); FRAC = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
f = None; hdr = None; seen = {}; 
for r in rows:
    if not r: continue
    if r[0] == "File Path" or r[0] == "File Name": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; iA = 2; iI = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or not r[iA].startswith("0x"): continue
    seen[r[iA]] = (f, int(r[0]), r[1].strip()[:90], int(r[iI] or 0))
hot = [v for v in seen.values() if v[3] >= FRAC * B]
print("static", len(seen), "=", len(seen) * 16 // 1024, "KB; hot (>= %.2f B)" % FRAC, len(hot), "=", len(hot) * 16 // 1024, "KB")
byfile = collections.Counter(); byline = collections.Counter(); src = {}
for fl, ln, s, i in hot: byfile[fl] += 1; byline[(fl, ln)] += 1; src[(fl, ln)] = s
for fl, c in byfile.most_common(): print(f"  {fl:30s} {c:5d} instr {c * 16 / 1024:5.1f} KB")
for (fl, ln), c in byline.most_common(N): print(f"{fl}:{ln:4d} {c:4d} | {src[(fl, ln)]}")
