#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, hot SASS regions by executed instructions, stall mix.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct"]
r = rows[2 + kidx]
print("kernel:", r[hdr.index("Kernel Name")][:70])
for k in keys:
    if k in hdr:
        print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
blocks, cur, h = [], None, None
for x in rows:
    if x and x[0] == "Address":
        h = x
        cur = []
        blocks.append(cur)
    elif cur is not None and h and len(x) == len(h):
        cur.append(x)
data = blocks[kidx]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
tot = sum(int(x[ia]) for x in data)
print("static instr %d, executed warp-instr %d" % (len(data), tot))
prev, start, acc, n, out = None, 0, 0, 0, []
for i, x in enumerate(data):
    c = int(x[ia])
    if prev is None or abs(c - prev) > 0.02 * max(c, prev, 1):
        if prev is not None:
            out.append((start, i - 1, prev, acc, n))
        start, acc, n = i, 0, 0
    prev = c
    acc += c
    n += 1
out.append((start, len(data) - 1, prev, acc, n))
for b in out:
    if b[3] > 0.004 * tot:
        ops = collections.Counter(data[j][isrc].split()[0].split(".")[0] if not data[j][isrc].startswith("@") else data[j][isrc].split()[1].split(".")[0] for j in range(b[0], b[1] + 1))
        print("  idx %4d-%4d x%-8d n=%4d share %5.1f%%  %s" % (b[0], b[1], b[2], b[4], 100.0 * b[3] / tot, dict(ops.most_common(6))))
cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tots = {h[i]: sum(int(x[i] or 0) for x in data) for i in cols}
s = sum(tots.values()) or 1
print("stalls:", {k: round(100.0 * v / s, 1) for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]})
