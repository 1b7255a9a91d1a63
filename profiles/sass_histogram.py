#!/usr/bin/env python
"""SASS opcode histogram per kernel of vgan_b200/lib/libvgan_gpu.so (cuobjdump -sass; runs without a GPU).
Usage: python profiles/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vgan_b200", "lib", "libvgan_gpu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print("# cuobjdump -sass %s" % os.path.relpath(lib, ROOT))
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("# architectures in the fat binary:", ", ".join(archs))
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
MARK = ("UBLKCP", "SYNCS", "REDG", "LDS", "LDG", "STG", "ATOMG", "DADD", "DFMA", "DMUL", "MUFU", "I2F", "VOTE", "SHFL", "BAR")
for k in sorted(hist):
    h = hist[k]
    if not h:
        continue
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = name.split("(")[0]
    tot = sum(h.values())
    groups = collections.Counter()
    for op, n in h.items():
        groups[op.split(".")[0]] += n
    marks = ["%s x%d" % (op, n) for op, n in sorted(h.items()) if any(op.startswith(mk) for mk in ("UBLKCP", "SYNCS", "REDG", "UTMA", "UTC", "HMMA", "ATOMG"))]
    print("\n%s  (%d instructions)" % (name, tot))
    print("  top:   " + ", ".join("%s %d" % (op, n) for op, n in groups.most_common(14)))
    if marks:
        print("  marks: " + ", ".join(marks))
