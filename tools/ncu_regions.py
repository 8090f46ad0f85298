#!/usr/bin/env python
"""Executed-instruction / stall-sample share of one kernel per SOURCE LINE, from an .ncu-rep captured with
--import-source on (ncu's own cuda,sass correlation; first launch that matches).
    python tools/ncu_regions.py REPORT.ncu-rep KERNEL_REGEX [min_pct]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
agg = collections.OrderedDict()
fname, hdr, seen_fn, nfn = "?", None, None, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        if seen_fn is None: seen_fn = r[1]
        elif r[1] != seen_fn: break
        continue
    if r and r[0] == "Line No":
        hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0]:   # a source line row (aggregated over its SASS)
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [r[1].strip(), 0, 0])
        a[1] += int(r[ci] or 0); a[2] += int(r[cs] or 0)
tot_i = sum(a[1] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"# {seen_fn[:90]}: {tot_i} warp-instr, {tot_s} samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if 100 * a[1] / tot_i >= min_pct or 100 * a[2] / tot_s >= min_pct:
        print(f"{100 * a[1] / tot_i:5.1f}%i {100 * a[2] / tot_s:5.1f}%s  {f}:{ln:<5d} {a[0][:120]}")
