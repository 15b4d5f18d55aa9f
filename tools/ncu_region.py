#!/usr/bin/env python
"""Stall-reason totals over a range of SASS indices. Usage: ncu_region.py rep lo hi"""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = None; data = []
for x in rows:
    if x and x[0] == "Address": h = x; continue
    if h and len(x) == len(h): data.append(x)
cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
isamp = h.index("# Samples"); iex = h.index("Instructions Executed")
tot = sum(int(x[isamp] or 0) for x in data)
sel = data[lo:hi + 1]
s = sum(int(x[isamp] or 0) for x in sel)
print("region %d-%d: %d samples of %d (%.1f%%), executed %d" % (lo, hi, s, tot, 100.0 * s / tot, sum(int(x[iex] or 0) for x in sel)))
agg = {h[c][6:]: sum(int(x[c] or 0) for x in sel) for c in cols}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    if v: print("  %-22s %6d  %5.1f%%" % (k, v, 100.0 * v / max(s, 1)))
