#!/usr/bin/env python
"""Executed warp-instructions per CUDA source line from an .ncu-rep (captured with --import-source on).
Usage: python tools/ncu_lines.py prof.ncu-rep [min_share_pct]"""
import csv
import io
import subprocess
import sys

def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = None
lines = []  # (file, line, src, inst, ops)
cur_file = None
for x in rows:
    if x and x[0] == "File Path":
        cur_file = x[1]
    elif x and x[0] == "Line No":
        h = x
        ie = h.index("Instructions Executed")
    elif h and len(x) == len(h):
        if x[0] != "":
            lines.append([cur_file, int(x[0]), x[1], num(x[ie]), {}])
        elif lines:
            op = x[3].split()
            if op and op[0].startswith("@"):
                op = op[1:]
            if op:
                k = op[0].split(".")[0]
                lines[-1][4][k] = lines[-1][4].get(k, 0) + num(x[ie])
tot = sum(l[3] for l in lines)
print("total executed warp-instr: %d" % tot)
for f, ln, src, n, ops in lines:
    if n >= thr / 100.0 * tot:
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        print("%5.1f%% %s:%d  %s\n        %s" % (100.0 * n / tot, f.split("/")[-1], ln, src.strip()[:100], {k: round(100.0 * v / tot, 2) for k, v in top}))
