#!/usr/bin/env python
"""Stall samples of one kernel in an .ncu-rep per CUDA source line (needs --import-source on).
usage: tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
H = rows[hi]
si, ie = H.index("# Samples"), H.index("Instructions Executed")
stall = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0
lines = [r for r in rows[hi + 1:] if r and r[0].isdigit() and len(r) > si]
tot = sum(num(r[si]) for r in lines)
print("samples", tot)
for r in lines:
    s = num(r[si])
    if 100 * s / tot >= min_pct:
        top = sorted(((num(r[c]), H[c]) for c in stall), reverse=True)[:3]
        print("%5s %6.2f%% exec=%-10s %-70s %s" % (r[0], 100 * s / tot, r[ie], r[1].strip()[:70], " ".join("%s=%d" % (n, v) for v, n in top if v)))
