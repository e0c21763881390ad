#!/usr/bin/env python
"""Aggregate ncu warp-stall samples per CUDA source line: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
agg, cur, ci, ei = {}, None, None, None
for x in csv.reader(out.splitlines()):
    if not x: continue
    if x[0] == "Line No":
        ci, ei = x.index("# Samples"), x.index("Instructions Executed"); continue
    if x[0] in ("File Path", "Function Name", "File Name"):
        fn = x[1] if len(x) > 1 else ""; continue
    if ci is None or len(x) <= max(ci, ei): continue
    if x[0] != "":
        try: cur = (fn[-40:], int(x[0]), x[1].strip()[:96])
        except ValueError: cur = None
    elif cur and x[ci].isdigit():
        a = agg.setdefault(cur, [0, 0]); a[0] += int(x[ci]); a[1] += int(x[ei]) if x[ei].isdigit() else 0
tot = sum(a[0] for a in agg.values())
print("total samples", tot, "total warp instrs", sum(a[1] for a in agg.values()))
for (f, l, src), (s, e) in sorted(agg.items(), key=lambda a: -a[1][0])[:topn]:
    print("%6d %5.1f%%  L%4d exec=%10d  %s" % (s, 100.0 * s / max(tot, 1), l, e, src))
