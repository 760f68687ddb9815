"""Summarise an .ncu-rep: key raw metrics per kernel + opcode / stall histograms from the source page.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] [units-per-launch]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else ""
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if kre and kre not in name:
        continue
    print("=====", name)
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]}")

args = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kre:
    args += ["--kernel-name", "regex:" + kre]
src = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
byop = collections.Counter(); stall = collections.Counter(); tot = 0
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        h = {k: i for i, k in enumerate(r)}
        cols = [k for k in r if k.startswith("stall_") and "Not Issued" not in k]
        continue
    if h is None or len(r) < len(h):
        continue
    try:
        n = int(r[h["Instructions Executed"]])
    except ValueError:
        continue
    op = r[h["Source"]].split()
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0]
    byop[o] += n
    tot += n
    for c in cols:
        stall[c] += int(r[h[c]])
print("total warp-instructions", tot, ("per unit %.1f" % (tot / units)) if units else "")
for o, n in byop.most_common(28):
    print(f"  {o:10s} {n:12d}" + (f"  {n / units:8.1f}/unit" if units else ""))
ts = sum(stall.values())
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in sorted(stall.items(), key=lambda x: -x[1]) if v))
