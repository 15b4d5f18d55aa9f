"""ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [TOP]: the SASS page of an ncu report (--import-source on) joined with nvdisasm's line table of the
same build, summed per source line: share of stall samples, share of warp instructions, lanes active, long-scoreboard and no-instruction share."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h, data = rows[hdr], rows[hdr + 1:]
nxt = [i for i, r in enumerate(data) if r and r[0] == "Kernel Name"]
if nxt:  # several launches of the kernel in the report: the first one
    data = data[:nxt[0]]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
ins = []
for f in sorted(os.listdir(tmp)):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in dis:
        continue
    on, cur = False, None
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
        if m:
            on = kern in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(2), cur))
assert len(ins) == len(data), (len(ins), len(data))
ci = {n: i for i, n in enumerate(h)}


def f(r, k):
    try:
        return float(r[ci[k]].replace(",", ""))
    except ValueError:
        return 0.0


keys = ["# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_no_inst", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"]
agg, tot = collections.defaultdict(lambda: [0.0] * 7), [0.0] * 7
for (txt, cur), r in zip(ins, data):
    for i, k in enumerate(keys):
        v = f(r, k)
        agg[cur][i] += v
        tot[i] += v
print("samples %d  warp inst %.3g  lanes %.2f  long_sb %.0f%%  no_inst %.0f%%" % (tot[0], tot[1], tot[2] / tot[1], 100 * tot[3] / tot[0], 100 * tot[4] / tot[0]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %5d  smp %5.1f%%  inst %5.1f%%  lanes %5.1f  long_sb %3.0f%%  no_inst %3.0f%%  smem wavefronts %5.1f%% (excess %5.1f%%)" %
          (k[0], k[1], 100 * v[0] / tot[0], 100 * v[1] / tot[1], v[2] / max(v[1], 1), 100 * v[3] / max(v[0], 1), 100 * v[4] / max(v[0], 1),
           100 * v[5] / max(tot[5], 1), 100 * v[6] / max(tot[5], 1)))
