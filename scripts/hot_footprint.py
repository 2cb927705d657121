"""Hot static code footprint per source line: static SASS counts per line from the object file (nvdisasm line info)
joined by source TEXT with the dynamic per-line instruction counts of an ncu source page
(`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`).
usage: hot_footprint.py obj.o kernel_pattern src.csv B [FRAC] [N]"""
import collections, csv, os, re, subprocess, sys, tempfile
obj, pat, srccsv, B = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4])
FRAC = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
N = int(sys.argv[6]) if len(sys.argv) > 6 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=d, capture_output=True, text=True).stdout.split("\n")
insec = False; cur = ("?", 0); cnt = collections.Counter()
for l in txt:
    if l.startswith("//--------------------- .text."): insec = pat in l; continue
    if l.startswith("//--------------------- "): insec = False; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): cnt[cur] += 1
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
srcs = {}
def text(f, ln):
    if f not in srcs:
        srcs[f] = None
        for base in ("predpreygrass_b200/csrc", "include"):
            p = os.path.join(root, base, f)
            if os.path.exists(p): srcs[f] = open(p).read().split("\n")
    s = srcs[f]
    return s[ln - 1].strip() if s and 0 < ln <= len(s) else ""
dyn = collections.Counter(); f = None; hdr = None
for r in csv.reader(open(srccsv)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): f = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= iI or not r[0].isdigit(): continue
    try: dyn[r[1].strip()] += int(r[iI])
    except ValueError: pass
tot = sum(cnt.values()); hot = []; unk = 0
for (fl, ln), c in cnt.items():
    t = text(fl, ln)
    if not t: unk += c; continue
    dn = dyn.get(t, 0)
    if dn >= FRAC * B * c: hot.append((c, fl, ln, t, dn))
print(f"static {tot} = {tot*16//1024} KB; no source text {unk}; hot (line executed >= {FRAC} B times per static instr): {sum(h[0] for h in hot)} = {sum(h[0] for h in hot)*16//1024} KB")
byf = collections.Counter()
for c, fl, ln, t, dn in hot: byf[fl] += c
for fl, c in byf.most_common(): print(f"  {fl:30s} {c:5d} = {c*16/1024:5.1f} KB")
for c, fl, ln, t, dn in sorted(hot, reverse=True)[:N]: print(f"{fl}:{ln:4d} static {c:4d} dyn/env {dn/B:6.1f} | {t[:100]}")
