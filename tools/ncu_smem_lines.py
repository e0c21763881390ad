#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line: ncu_smem_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
agg, cur, wi, ii, ei = {}, None, None, None, None
for x in csv.reader(out.splitlines()):
    if not x: continue
    if x[0] == "Line No":
        wi, ii, ei = x.index("L1 Wavefronts Shared"), x.index("L1 Wavefronts Shared Ideal"), x.index("Instructions Executed"); continue
    if x[0] in ("File Path", "Function Name", "File Name"):
        fn = x[1] if len(x) > 1 else ""; continue
    if wi is None or len(x) <= max(wi, ii): continue
    if x[0] != "":
        try: cur = (int(x[0]), x[1].strip()[:100])
        except ValueError: cur = None
    elif cur and x[wi].isdigit():
        a = agg.setdefault(cur, [0, 0, 0]); a[0] += int(x[wi]); a[1] += int(x[ii]) if x[ii].isdigit() else 0
        a[2] += int(x[ei]) if x[ei].isdigit() else 0
tot = sum(a[0] for a in agg.values())
print("total shared wavefronts", tot, "ideal", sum(a[1] for a in agg.values()))
for (l, src), (w, i, e) in sorted(agg.items(), key=lambda a: -a[1][0])[:topn]:
    print("%10d %5.1f%% ideal=%10d inst=%9d L%4d  %s" % (w, 100.0 * w / max(tot, 1), i, e, l, src))
