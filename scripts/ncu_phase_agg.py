import csv, sys, re
# usage: ncu_phase_agg.py src_both.csv file.cu  -- aggregates instructions/samples per "phase" using marker comments in the .cu
rows=list(csv.reader(open(sys.argv[1])))
src=open(sys.argv[2]).read().split('\n')
marks=[]  # (line, name)
for i,l in enumerate(src,1):
    m=re.search(r'//\s*PHASE:\s*(.*)',l)
    if m: marks.append((i,m.group(1).strip()))
def phase(line):
    name='(pre)'
    for ln,nm in marks:
        if line>=ln: name=nm
    return name
agg={}; f=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': f=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; n=len(hdr); iS=hdr.index('# Samples')-n; iI=hdr.index('Instructions Executed')-n; continue
    if len(r)<10 or r[0]=='': continue
    try: ln=int(r[0]); s=float(r[iS] or 0); i=float(r[iI] or 0)
    except ValueError: continue
    key=phase(ln) if f==sys.argv[2].split('/')[-1] else 'other:'+f
    a=agg.setdefault(key,[0,0]); a[0]+=s; a[1]+=i
ts=sum(a[0] for a in agg.values()); ti=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:40s} samples {a[0]:7.0f} ({100*a[0]/ts:4.1f}%)  instr {a[1]:10.0f} ({100*a[1]/ti:4.1f}%)")
