#!/usr/bin/env python
"""Key per-launch metrics from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv
import sys

r = list(csv.reader(sys.stdin))
h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
idx = [h.index(w) for w in want if w in h]
units = r[1]
print(",".join(f"{h[i]} [{units[i]}]" for i in idx))
for row in r[2:]:
    print(",".join(row[i].split("(")[0] if h[i] == "Kernel Name" else row[i] for i in idx))
