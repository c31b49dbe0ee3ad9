#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print(f"{k[:40]:40s} n={cnt[k]:4d} total={v:10.1f} us avg={v / cnt[k]:8.1f} share={100 * v / T:5.1f}%")
