#!/usr/bin/env python3
"""Summarise an ncu launch list (gpu__time_duration per launch): usage scripts/launch_summary.py <csv> [kernel-substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
head = rows[hi]
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in rows[hi + 1:]:
    if len(r) < len(head):
        continue
    name = r[head.index('Kernel Name')].split('(')[0]
    v = float(r[head.index('Metric Value')].replace(',', ''))
    unit = r[head.index('Metric Unit')]
    v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
    agg[name][0] += 1
    agg[name][1] += v
    seq.append((name, v))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} {v[0]:4d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%")
print(f"total {tot:.1f} us, {len(seq)} launches")
if len(sys.argv) > 2:
    print([round(v, 1) for n, v in seq if sys.argv[2] in n])
