#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --set full` raw CSV: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a
kernel, averaged over the captured launches, with the launch state it was captured in.  bench.py reads it for `roofline.traffic`.
usage: ncu_traffic.py <raw.csv> <kernel substring> <json key> <source note> [key=value ...]   (extra key=value pairs are stored as ints)"""
import csv, json, os, subprocess, sys

raw, kern, key, note = sys.argv[1:5]
extra = dict(kv.split("=") for kv in sys.argv[5:])
r = list(csv.reader(open(raw)))
h, units = r[0], r[1]
def col(name):
    i = h.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1.0)
    return [float(row[i].replace(",", "")) * scale for row in r[2:] if kern in row[h.index("Kernel Name")]]
rd, wr, us = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
out = {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "launches": len(rd), "dram_bytes_read": rd, "dram_bytes_write": wr,
       "ncu_duration_us": us, "source": note,
       "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()}
out.update({k: int(v) for k, v in extra.items()})
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_traffic.json")
d = json.load(open(p)) if os.path.exists(p) else {}
d[key] = out
json.dump(d, open(p, "w"), indent=1)
print(key, out["dram_bytes_per_launch"], "bytes/launch over", len(rd), "launches")
