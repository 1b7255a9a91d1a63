#!/usr/bin/env python
"""Top stall locations of a kernel from an .ncu-rep captured with --set full --import-source on (SASS view)."""
import csv, subprocess, sys
rep = sys.argv[1]
reads = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if "Instructions Executed" in r:
        h, start = r, i
        break
ie, si, ai, sm, te = (h.index(x) for x in ("Instructions Executed", "Source", "Address", "# Samples", "Thread Instructions Executed"))
stall_cols = [k for k, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
ins = []
for r in rows[start + 1:]:
    try:
        ins.append((int(r[ai], 16), r[si].strip(), int(r[ie]), int(r[sm]), int(r[te]), r))
    except Exception:
        pass
base = ins[0][0]
ts = sum(x[3] for x in ins)
print("instructions/read %.0f   samples %d" % (sum(x[2] for x in ins) / reads, ts))
idx = {x[0]: k for k, x in enumerate(ins)}
for a, s, n, smp, t, r in sorted(ins, key=lambda x: -x[3])[:ntop]:
    why = sorted(((int(r[k]), h[k][6:]) for k in stall_cols if r[k] not in ("", "0")), reverse=True)[:2]
    k = idx[a]
    print("%5x %5.1f%% lanes=%4.1f n/read=%5.2f %-44s %s" % (a - base, 100 * smp / ts, t / max(n, 1), n / reads, s[:44], why))
    for x in ins[max(0, k - 3):k]:
        print("          %-60s" % x[1][:60])
