#!/usr/bin/env python
"""Kernel shares of the LAST `nsteps` frames in an `ncu --metrics gpu__time_duration.sum --csv` launch list (frames are delimited by
k_classify launches)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
top = int(sys.argv[3]) if len(sys.argv) > 3 else 24
names = [r['Kernel Name'].split('(')[0] for r in rows]
idx = [i for i, n in enumerate(names) if n == 'k_classify']
start = idx[-nsteps]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[start:]:
    n = r['Kernel Name'].split('(')[0]; v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    tot[n] += v; cnt[n] += 1
T = sum(tot.values())
print(f"{nsteps} frames: {T:.1f} us in {sum(cnt.values())} launches = {T / nsteps / 1000:.2f} ms per frame")
for k, v in tot.most_common(top):
    print(f"{k[:44]:44s} n={cnt[k]:4d} total={v:9.1f} avg={v / cnt[k]:8.1f} share={100 * v / T:5.1f}%")
