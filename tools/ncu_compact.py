#!/usr/bin/env python
"""One compact line of key ncu metrics per captured launch, from a tools/ncu_summary.py digest.
    python tools/ncu_compact.py digest.txt"""
import sys
KEYS = {"gpu__time_duration.sum": "t_us", "dram__bytes_read.sum": "dR", "dram__bytes_write.sum": "dW",
        "lts__t_sector_hit_rate.pct": "L2hit", "l1tex__t_sector_hit_rate.pct": "L1hit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
        "smsp__inst_executed.sum": "inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
        "launch__registers_per_thread": "regs", "launch__waves_per_multiprocessor": "waves",
        "stalled_long_scoreboard_per_issue": "LSB", "stalled_barrier_per": "BAR", "stalled_short_scoreboard_per": "SSB",
        "stalled_lg_throttle_per": "LG", "stalled_mio_throttle_per": "MIO", "stalled_math_pipe": "MATH",
        "stalled_wait_per": "WAIT", "stalled_not_selected": "NSEL", "stalled_no_instruction": "NOI",
        "stalled_branch_resolving": "BR"}
for b in open(sys.argv[1]).read().split("== ")[1:]:
    lines = b.split("\n")
    out = []
    for l in lines[1:]:
        for k, s in KEYS.items():
            if k in l:
                p = l.split()
                try:
                    vs = "%.3g" % float(p[1])
                except ValueError:
                    vs = p[1]
                if s in ("dR", "dW") and len(p) > 2:
                    vs += p[2][0]
                out.append(s + "=" + vs)
                break
    print(lines[0][:70])
    print("    " + " ".join(out))
