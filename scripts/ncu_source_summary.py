#!/usr/bin/env python
"""Condense an `ncu --page source --csv` dump (per-SASS-instruction stall samples)
of one capture: stall-reason shares, executed-opcode mix, hottest instructions.
  python scripts/ncu_source_summary.py <source.csv> [kernel index] [top n]"""
import csv,sys,collections
f=sys.argv[1]; which=int(sys.argv[2]) if len(sys.argv)>2 else 0
rows=list(csv.reader(open(f)))
# split into kernels
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
starts.append(len(rows))
a,b=starts[which],starts[which+1]
hdr=rows[a+1]; data=rows[a+2:b]
ix={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot=collections.Counter(); n=0
samp=ix["# Samples"]
recs=[]
for r in data:
    if len(r)<len(hdr): continue
    try: s=int(r[samp])
    except: continue
    n+=s
    for c in stall_cols:
        try: tot[c]+=int(r[ix[c]])
        except: pass
    recs.append(r)
print("instructions",len(recs),"samples",n)
for c,v in tot.most_common(10): print(f"  {c:28s} {v:8d} {100*v/max(n,1):5.1f}%")
# executed instruction totals by opcode class
ex=collections.Counter(); exs=collections.Counter()
for r in recs:
    op=r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
    if op.startswith("@"): op=r[ix["Source"]].split()[1]
    op=op.split(".")[0]
    ex[op]+=int(r[ix["Instructions Executed"]]); exs[op]+=int(r[samp])
te=sum(ex.values())
print("executed warp instrs",te)
for op,v in ex.most_common(16): print(f"  {op:10s} exec {100*v/te:5.1f}%  samples {100*exs[op]/max(n,1):5.1f}%")
# top 25 instructions by samples
print("top instrs by samples")
top=sorted(recs,key=lambda r:-int(r[samp]))[:int(sys.argv[3]) if len(sys.argv)>3 else 25]
for r in top:
    st=sorted(((int(r[ix[c]]),c) for c in stall_cols if r[ix[c]] not in ("","0")),reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {int(r[samp]):6d} {r[ix['Source']][:60]:60s} {st}")
