#!/usr/bin/env python
"""Top source lines of one kernel in an .ncu-rep by executed instructions / stall samples.
    python tools/ncu_source_top.py REPORT.ncu-rep KERNEL_REGEX [N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several tables may follow each other (one per launch / per file); parse each header block
i = 0
agg = {}
while i < len(rows):
    r = rows[i]
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
        cs = hdr.index("# Samples") if "# Samples" in hdr else None
        i += 1
        while i < len(rows) and len(rows[i]) == len(hdr):
            q = rows[i]
            try:
                ln = int(q[0])
            except ValueError:
                break
            ins = int(q[ci] or 0) if ci is not None else 0
            smp = int(q[cs] or 0) if cs is not None else 0
            a = agg.setdefault((fname, ln), [q[1], 0, 0])
            a[1] += ins
            a[2] += smp
            i += 1
        continue
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    i += 1
tot_i = sum(a[1] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total instructions {tot_i}  samples {tot_s}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"{100 * a[1] / tot_i:5.1f}%i {100 * a[2] / tot_s:5.1f}%s  {f}:{ln:<5d} {a[0].strip()[:110]}")
