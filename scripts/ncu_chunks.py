"""Per-region view of an ncu source page (ncu -i X.ncu-rep --page source --csv > file): samples, stall mix and opcode mix
per chunk of SASS instructions, plus totals.  usage: python scripts/ncu_chunks.py file.csv [chunk]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r[idx["# Samples"]]) for r in data)
ex_tot = sum(num(r[idx["Instructions Executed"]]) for r in data)
print(f"{len(data)} SASS instructions ({len(data) * 16} bytes), {tot:.0f} samples, {ex_tot / 1e6:.2f} M warp-instructions executed")
agg = collections.Counter()
for r in data:
    for h in stalls:
        agg[h] += num(r[idx[h]])
print("  ".join(f"{h[6:]}:{100 * v / tot:.1f}%" for h, v in agg.most_common(10)))
for c in range(0, len(data), chunk):
    seg = data[c:c + chunk]
    s = sum(num(r[idx["# Samples"]]) for r in seg)
    ex = sum(num(r[idx["Instructions Executed"]]) for r in seg)
    if ex < 1e5:
        continue
    ag, ops = collections.Counter(), collections.Counter()
    for r in seg:
        for h in stalls:
            ag[h] += num(r[idx[h]])
        src = r[idx["Source"]].strip()
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        ops[op.split(".")[0]] += 1
    top = ", ".join(f"{h[6:]}:{int(v)}" for h, v in ag.most_common(3))
    topo = ", ".join(f"{o}:{n}" for o, n in ops.most_common(4))
    print(f"{c:5d} ex/instr {ex / len(seg) / 1e3:7.1f}k samples {int(s):5d} {100 * s / tot:5.1f}% | {top:45s} | {topo}")
