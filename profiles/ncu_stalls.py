#!/usr/bin/env python
"""Per-launch headline metrics + top warp-stall reasons from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv, sys
r = list(csv.reader(sys.stdin)); h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum"]
stall = [x for x in h if x.startswith("smsp__average_warp") and "issue_stalled" in x and x.endswith("per_issue_active.ratio")]
for row in r[2:]:
    print({w.split(".")[0].replace("smsp__", "").replace("launch__", ""): row[h.index(w)][:48] for w in want if w in h})
    s = sorted(((float(row[h.index(x)].replace(",", "")), x) for x in stall if row[h.index(x)] not in ("", "n/a")), reverse=True)[:6]
    for v, x in s:
        print("    %6.2f %s" % (v, x.replace("smsp__average_", "").replace("_per_issue_active.ratio", "")))
