#!/usr/bin/env python
"""Shared-memory wavefronts / excessive wavefronts (bank conflicts) per SASS instruction and CUDA source line.
Usage: python tools/ncu_smem.py prof.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, cur, out = None, None, []
for x in rows:
    if x and x[0] == "Line No":
        h = x
        iw, ix, ie = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Excessive"), h.index("Instructions Executed")
    elif h and len(x) == len(h):
        if x[0] != "":
            cur = (int(x[0]), x[1].strip()[:90])
        else:
            try:
                w, e, n = int(x[iw] or 0), int(x[ix] or 0), int(x[ie] or 0)
            except ValueError:
                continue
            if w:
                out.append((e, w, n, x[3].strip()[:60], cur))
tw, te = sum(o[1] for o in out), sum(o[0] for o in out)
print("shared wavefronts %d, excessive %d (%.1f%%)" % (tw, te, 100.0 * te / max(tw, 1)))
for e, w, n, sass, cur in sorted(out, key=lambda o: -o[0])[:top]:
    cur = cur or (0, "")
    print("exc %8d  wf %8d  exec %7d  %-60s | %d: %s" % (e, w, n, sass, cur[0], cur[1]))
