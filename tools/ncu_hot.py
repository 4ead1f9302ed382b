#!/usr/bin/env python
"""Hot instructions of one kernel in an .ncu-rep (source page): every instruction above a share of the
stall samples, in program order, plus a few kernel-level figures.  usage: tools/ncu_hot.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, V = rows[0], rows[2]
for key in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
    for h, u, v in zip(H, rows[1], V):
        if h == key:
            print("%-70s %s %s" % (h, v, u))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
H = rows[1]
data = [r for r in rows[2:] if len(r) > 5]
si, so, ie = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[si] or 0) for r in data)
print("samples %d, instructions %d" % (tot, len(data)))
for i, r in enumerate(data):
    s = float(r[si] or 0)
    if 100 * s / tot >= min_pct:
        top = sorted(((float(r[c] or 0), H[c]) for c in stall_cols), reverse=True)[:2]
        print("%5d %6.2f%% exec=%-8s %-60s %s" % (i, 100 * s / tot, r[ie], r[so].strip()[:60],
                                                   " ".join("%s=%d" % (n, v) for v, n in top if v)))
