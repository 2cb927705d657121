import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
agg={}
f=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': f=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; n=len(hdr); iS=hdr.index('# Samples')-n; iI=hdr.index('Instructions Executed')-n; continue
    if len(r)<10 or r[0]=='': continue
    try:
        key=(f,int(r[0])); s=float(r[iS] or 0); i=float(r[iI] or 0)
    except ValueError: continue
    a=agg.setdefault(key,[r[1],0,0]); a[1]+=s; a[2]+=i
tot_s=sum(a[1] for a in agg.values()); tot_i=sum(a[2] for a in agg.values())
print('total samples',tot_s,'instr',tot_i)
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
print('--- by samples')
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:N]:
    print(f"{k[0]}:{k[1]:4d} s={a[1]:6.0f} ({100*a[1]/tot_s:4.1f}%) i={a[2]:9.0f} ({100*a[2]/tot_i:4.1f}%) {a[0][:100]}")
print('--- by instr')
for k,a in sorted(agg.items(), key=lambda x:-x[1][2])[:N]:
    print(f"{k[0]}:{k[1]:4d} s={a[1]:6.0f} ({100*a[1]/tot_s:4.1f}%) i={a[2]:9.0f} ({100*a[2]/tot_i:4.1f}%) {a[0][:100]}")
