#!/usr/bin/env python
"""Stall samples of a kernel grouped by (stall reason, opcode of the waiting instruction) and instruction mix,
from an .ncu-rep captured with --set full --import-source on."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
reads = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if "Instructions Executed" in r:
        h, start = r, i
        break
ie, si, sm = (h.index(x) for x in ("Instructions Executed", "Source", "# Samples"))
stall_cols = [k for k, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
by = collections.Counter(); mix = collections.Counter(); tot = 0
for r in rows[start + 1:]:
    try:
        n = int(r[ie]); s = int(r[sm])
    except Exception:
        continue
    op = r[si].strip().split()
    op = (op[1] if op[0].startswith("@") else op[0]).split(".")[0] if op else "?"
    mix[op] += n
    for k in stall_cols:
        if r[k] not in ("", "0"):
            by[(h[k][6:], op)] += int(r[k]); tot += int(r[k])
print("stall samples by (reason, waiting opcode):")
for (why, op), v in by.most_common(18):
    print("  %5.1f%%  %-18s %s" % (100.0 * v / tot, why, op))
print("warp instructions per read by opcode:")
for op, n in mix.most_common(22):
    print("  %7.2f  %s" % (n / reads, op))
