#!/usr/bin/env python
"""Top SASS instructions by stall samples, with their dominant stall reasons.  Usage: ncu_stalls.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = None; data = []
for x in rows:
    if x and x[0] == "Address": h = x; continue
    if h and len(x) == len(h): data.append(x)
isamp = h.index("# Samples"); isrc = h.index("Source"); iex = h.index("Instructions Executed")
cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(x[isamp] or 0) for x in data)
print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top]
for i in sorted(order):
    x = data[i]
    st = sorted(((int(x[c] or 0), h[c][6:]) for c in cols), reverse=True)[:3]
    print("%5d %5.1f%% exec %9s  %-58s %s" % (i, 100.0 * int(x[isamp] or 0) / tot, x[iex], x[isrc].strip()[:58], ", ".join("%s %d" % (n, v) for v, n in st if v)))
