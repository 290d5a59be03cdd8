#!/usr/bin/env python
"""Selected raw metrics per kernel of an ncu report:  python tools/ncu_summary.py rep.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    rec = {}
    for w in WANT:
        if w in d:
            rec[w] = d[w] + (" " + u[w] if u.get(w) else "")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(d[k]) >= 0.15:
                    rec["stall cycles per issue: " + k.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")] = d[k]
            except ValueError:
                pass
    out.append(rec)
    for k, v in rec.items():
        print(f"{k:70s} {v}")
    print()
if len(sys.argv) > 2:
    keys = []
    for rec in out:
        for k in rec:
            if k not in keys:
                keys.append(k)
    with open(sys.argv[2], "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(keys)
        for rec in out:
            wr.writerow([rec.get(k, "") for k in keys])
