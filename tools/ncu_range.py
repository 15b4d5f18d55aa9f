#!/usr/bin/env python
"""Stall samples summed over ranges of SASS instruction indices.  Usage: ncu_range.py rep lo-hi [lo-hi ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = None; data = []
for x in rows:
    if x and x[0] == "Address": h = x; continue
    if h and len(x) == len(h): data.append(x)
isamp = h.index("# Samples"); iex = h.index("Instructions Executed")
cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(x[isamp] or 0) for x in data)
for r in sys.argv[2:]:
    lo, hi = map(int, r.split("-"))
    sel = data[lo:hi + 1]
    n = sum(int(x[isamp] or 0) for x in sel)
    ex = sum(int(x[iex] or 0) for x in sel)
    st = sorted(((sum(int(x[c] or 0) for x in sel), h[c][6:]) for c in cols), reverse=True)[:7]
    print("%s: %.1f%% of samples, %d warp-instr; %s" % (r, 100.0 * n / tot, ex, ", ".join("%s %.1f%%" % (k, 100.0 * v / max(n, 1)) for v, k in st)))
