#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launch_summary.py <file.csv>"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    tot[r[4]][0] += float(r[-1].replace(",", "")) / 1e3
    tot[r[4]][1] += 1
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us total  n={n:3d}  avg {us / n:9.1f} us  {name[:110]}")
