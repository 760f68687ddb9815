"""Bucket a kernel's ncu source page by SASS position: sample share, executed instructions, opcode mix, top stalls.
   python scripts/ncu_buckets.py rep kernel-regex [bucket=60] [units]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 60
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kre],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
h=None; data=[]
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        h={k:i for i,k in enumerate(r)}; cols=[k for k in r if k.startswith("stall_") and "Not Issued" not in k]; continue
    if h is None or len(r)<len(h): continue
    try: n=int(r[h["Instructions Executed"]])
    except: continue
    data.append(r)
tot=sum(int(r[h["# Samples"]]) for r in data)
print("instrs",len(data),"samples",tot)
for b in range(0,len(data),B):
    chunk=data[b:b+B]
    s=sum(int(r[h["# Samples"]]) for r in chunk)
    ex=sum(int(r[h["Instructions Executed"]]) for r in chunk)
    ops={}
    for r in chunk:
        o=r[h["Source"]].split(); o=(o[1] if o[0].startswith("@") else o[0]).split(".")[0]; ops[o]=ops.get(o,0)+1
    st={}
    for r in chunk:
        for c in cols: st[c[6:]]=st.get(c[6:],0)+int(r[h[c]])
    top=sorted(st.items(),key=lambda x:-x[1])[:3]
    if s*100/tot>0.4:
        print(f"{b:5d} {100*s/tot:5.1f}% ex {ex/units:9.1f} ", ", ".join(f"{k}:{v}" for k,v in sorted(ops.items(),key=lambda x:-x[1])[:4]), " | ", ", ".join(f"{k} {100*v/max(1,s):.0f}%" for k,v in top))
