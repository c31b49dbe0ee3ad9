#!/usr/bin/env python
"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name K --launch-count 1` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[ismp]) for r in data)
print("total samples", tot, "SASS instr", len(data), "warp-instr executed", sum(int(r[iex]) for r in data))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 500
for b in range(0, len(data), B):
    s = sum(int(r[ismp]) for r in data[b:b + B])
    ex = sum(int(r[iex]) for r in data[b:b + B])
    print(f"  sass[{b:6d}..] samples {s:7d} {100 * s / max(tot, 1):5.1f}%  executed {ex}")
for r in sorted(data, key=lambda r: -int(r[ismp]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"  {r[ismp]:>6s} {r[iex]:>9s} @{data.index(r):6d} {r[isrc][:100]}")
