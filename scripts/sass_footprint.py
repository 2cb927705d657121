"""static code footprint per source line of one kernel, from the object file (no GPU needed)
usage: python scripts/sass_footprint.py predpreygrass_b200/csrc/ppg_eco.o step_eco_kernelILi1EhLb1E [N]"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=d, capture_output=True, text=True).stdout.split("\n")
insec = False; cur = ("?", 0); cnt = collections.Counter(); tot = 0
for l in txt:
    if l.startswith("//--------------------- .text."):
        insec = pat in l; continue
    if l.startswith("//--------------------- "):
        insec = False; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[cur] += 1; tot += 1
print("instructions", tot, "=", tot * 16 // 1024, "KB")
byf = collections.Counter()
for (f, ln), c in cnt.items(): byf[f] += c
for f, c in byf.most_common(): print(f"{f:32s} {c}")
src = {}
for (f, ln), c in cnt.most_common(N):
    if f not in src:
        for base in ("predpreygrass_b200/csrc", "include"):
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", base, f)
            if os.path.exists(p): src[f] = open(p).read().split("\n")
    line = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:110] if f in src else ""
    print(f"{f}:{ln} {c} | {line}")
