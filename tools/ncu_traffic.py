#!/usr/bin/env python
"""DRAM bytes per launch of every kernel in an .ncu-rep -> profiles/traffic_r01.json (read by bench.py).
    python tools/ncu_traffic.py REPORT.ncu-rep SHAPE FRAMES_PER_LAUNCH SOURCE_NOTE > profiles/traffic_r01.json"""
import csv
import io
import json
import subprocess
import sys

rep, shape, frames, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[col[k]].replace(",", "")) * scale[units[col[k]]]
    out[name] = int(tot)   # the last launch of a kernel wins (all launches of a step are alike)
print(json.dumps({"source": note, "shape": shape, "frames_per_launch": frames, "dram_bytes_per_launch": out,
                  "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch; data still resident in L2 "
                          "at kernel end is not counted"}, indent=1))
