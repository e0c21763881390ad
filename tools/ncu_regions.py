#!/usr/bin/env python
"""Instruction and stall-sample budget per source-line range: ncu_regions.py rep 'name:lo-hi' ..."""
import csv, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
ci = ei = None; cur = None; fn = ""
acc = {n: [0, 0] for n, _, _ in regions}; other = [0, 0]
for x in csv.reader(out.splitlines()):
    if not x: continue
    if x[0] == "Line No": ci, ei = x.index("# Samples"), x.index("Instructions Executed"); continue
    if x[0] in ("File Path", "File Name"): fn = x[1] if len(x) > 1 else ""; continue
    if x[0] == "Function Name": continue
    if ci is None or len(x) <= max(ci, ei): continue
    if x[0] != "":
        try: cur = (fn, int(x[0]))
        except ValueError: cur = None
    elif cur and x[ci].isdigit():
        s, e = int(x[ci]), int(x[ei]) if x[ei].isdigit() else 0
        tgt = other
        if cur[0].endswith("fpx_kernels.cu"):
            for n, lo, hi in regions:
                if lo <= cur[1] <= hi: tgt = acc[n]; break
        tgt[0] += s; tgt[1] += e
tot_s = sum(v[0] for v in acc.values()) + other[0]; tot_e = sum(v[1] for v in acc.values()) + other[1]
for n, _, _ in regions + [("other(headers/inlined)", 0, 0)]:
    v = acc.get(n, other)
    print("%-28s samples %7d %5.1f%%   warp-instrs %12d %5.1f%%" % (n, v[0], 100.0 * v[0] / tot_s, v[1], 100.0 * v[1] / tot_e))
print("total warp-instrs", tot_e)
