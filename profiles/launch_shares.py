#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total time, share.
usage: launch_shares.py <launches.csv> [header comment]"""
import csv, sys, collections
r = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, row in enumerate(r) if "Kernel Name" in row)
hd = r[h]; ki = hd.index("Kernel Name"); vi = hd.index("Metric Value"); ui = hd.index("Metric Unit")
agg = collections.OrderedDict()
for row in r[h + 1:]:
    if len(row) <= vi: continue
    t = float(row[vi].replace(",", ""))
    unit = row[ui]
    us = t / 1000.0 if unit in ("ns", "nsecond") else (t * 1000.0 if unit in ("ms", "msecond") else t)
    name = row[ki].split("(")[0].strip()
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2: print("# " + sys.argv[2])
print("# cold-cache, serialised: compare SHARES with bench.py's `kernels` (CUDA events), not absolutes")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:90]:90s} n={n:4d} total_us={us:10.1f} share={us / tot:6.3f} avg_us={us / n:8.1f}")
