#!/usr/bin/env python
"""Print the handful of ncu metrics we track for a kernel from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
reads = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
d = {}
for i, x in enumerate(h):
    if x in want or ("issue_stalled" in x and "per_issue_active" in x):
        d[x] = (u[i], v[i])
for k in want:
    if k in d:
        print("%-70s %-14s %s" % (k, d[k][1], d[k][0]))
st = sorted(((float(d[k][1]), k) for k in d if "issue_stalled" in k), reverse=True)[:6]
for val, k in st:
    print("  stall %-55s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], val))
if reads:
    print("warp-instructions per read: %.0f" % (float(d["smsp__inst_executed.sum"][1]) / reads))
    print("shared wavefronts per read: %.0f" % (float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][1]) / reads))
