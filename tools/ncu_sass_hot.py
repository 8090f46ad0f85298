#!/usr/bin/env python
"""Hot SASS of one kernel in an .ncu-rep: executed-instruction and stall-sample share per instruction,
grouped into basic regions.   python tools/ncu_sass_hot.py REPORT.ncu-rep KERNEL_REGEX [min_pct]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break  # first launch only
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
tot_i = sum(int(r[ci] or 0) for r in data) or 1
tot_s = sum(int(r[cs] or 0) for r in data) or 1
print(f"# {kern}: {len(data)} SASS instructions, executed {tot_i} warp-instr, {tot_s} samples")
base = int(data[0][0], 16)
for r in data:
    pi, ps = 100 * int(r[ci] or 0) / tot_i, 100 * int(r[cs] or 0) / tot_s
    if pi >= min_pct or ps >= min_pct:
        print(f"{int(r[0], 16) - base:6x} {pi:5.2f}%i {ps:5.2f}%s thr={r[ct]:>5s}  {r[1].strip()[:90]}")
