#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python scripts/launch_summary.py file.csv [...]"""
import collections
import csv
import sys

for f in sys.argv[1:]:
    tot, cnt = collections.Counter(), collections.Counter()
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    for r in rows:
        name = r[4].split("(")[0][:60]
        tot[name] += float(r[-1])
        cnt[name] += 1
    s = sum(tot.values()) or 1.0
    print(f)
    for k, v in tot.most_common(10):
        print(f"  {v / 1e6:9.3f} ms {100 * v / s:5.1f}%  x{cnt[k]:4d}  avg {v / cnt[k] / 1e3:8.1f} us  {k}")
