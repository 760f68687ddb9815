"""Per-kernel SASS opcode census of the built library: which Blackwell-specific instructions each kernel really contains
(B200_PROFILING.md, "What proves a Blackwell-native kernel").

    python scripts/sass_census.py [path/to/libkoemorph_b200.so] > profiles/sass_census_rNN.txt
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "koemorph_b200", "csrc", "libkoemorph_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "USETMAXREG", "HMMA",
       "FFMA2", "FADD2", "FMUL2", "FFMA", "LDG", "STG", "LDS", "STS", "LDGSTS", "SHFL", "REDUX", "MUFU", "BAR", "ATOM", "RED", "CCTL"]
fn, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[fn] = collections.Counter()
        total[fn] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        counts[fn][m.group(1)] += 1
        total[fn] += 1
print(f"# SASS census of {os.path.basename(lib)} (static instruction counts per kernel; cuobjdump -sass)")
for fn, c in counts.items():
    keys = ", ".join(f"{k} {c[k]}" for k in KEY if c[k])
    print(f"{fn}\n    {total[fn]} instructions ({total[fn] * 16} bytes): {keys}")
