"""Split a kernel's ncu source page into regions at BAR.SYNC and print per-region executed instructions, sample share
and stall mix.   python scripts/ncu_regions.py rep kernel-regex"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; regions = []; cur = None
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        h = {k: i for i, k in enumerate(r)}
        cols = [k for k in r if k.startswith("stall_") and "Not Issued" not in k]
        cur = dict(name="start", inst=0, samples=0, stall=collections.Counter(), ops=collections.Counter(), n=0)
        regions.append(cur)
        continue
    if h is None or len(r) < len(h):
        continue
    try:
        n = int(r[h["Instructions Executed"]])
    except ValueError:
        continue
    s = r[h["Source"]]
    cur["inst"] += n; cur["n"] += 1
    cur["samples"] += int(r[h["# Samples"]])
    op = s.split(); o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    cur["ops"][o] += n
    for c in cols:
        cur["stall"][c[6:]] += int(r[h[c]])
    if "BAR.SYNC" in s or "EXIT" in s.split()[-2:] or s.strip().startswith("EXIT"):
        cur = dict(name=f"after {s.strip()[:24]} @{r[h['Address']][-5:]}", inst=0, samples=0, stall=collections.Counter(),
                   ops=collections.Counter(), n=0)
        regions.append(cur)
tot = sum(x["samples"] for x in regions) or 1
for x in regions:
    if x["inst"] == 0 and x["samples"] == 0:
        continue
    st = ", ".join(f"{k} {100*v/max(1,sum(x['stall'].values())):.0f}%" for k, v in x["stall"].most_common(5))
    ops = ", ".join(f"{k} {v}" for k, v in x["ops"].most_common(6))
    print(f"{x['name']:40s} static {x['n']:5d} executed {x['inst']:11d} samples {100*x['samples']/tot:5.1f}%  [{st}]\n      {ops}")
