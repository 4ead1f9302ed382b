#!/usr/bin/env python
"""Turn gpurun_out/launches.csv (+ an optional .ncu-rep) into a small markdown summary under profiles/.
usage: tools/summarize_ncu.py <tag> [launches.csv] [report.ncu-rep]"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
reps = sys.argv[3:]
out = ["# ncu summary %s" % tag, ""]
rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: "
        "compare shares)", "", "| kernel | launches | mean ms | share |", "|---|---:|---:|---:|"]
for k, v in agg.items():
    out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    Hh, Uu = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    for row in rr[2:3]:
        name = row[Hh.index("Kernel Name")] if "Kernel Name" in Hh else "?"
        out += ["", "## `ncu --set full` of `%s`" % name.split("(")[0], "", "| metric | unit | value |", "|---|---|---:|"]
        for h, u, v in zip(Hh, Uu, row):
            if h in want:
                out.append("| %s | %s | %s |" % (h, u, v))
open("profiles/%s.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
