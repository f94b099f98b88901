#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum (usage: scripts/dram_summary.py <csv>)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
head = rows[hi]
ki, mi, vi, ui = (head.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
per, order = {}, []
for r in rows[hi + 1:]:
    if len(r) < len(head):
        continue
    if r[0] not in per:
        per[r[0]] = {'name': r[ki].split('(')[0]}
        order.append(r[0])
    v, u = float(r[vi].replace(',', '')), r[ui]
    if r[mi].startswith('gpu__time'):
        per[r[0]]['us'] = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    else:
        per[r[0]][r[mi]] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in order:
    p, a = per[i], agg[per[i]['name']]
    a[0] += 1
    a[1] += p['us']
    a[2] += p.get('dram__bytes_read.sum', 0)
    a[3] += p.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:60]:60s} {a[0]:4d} {a[1]:9.1f} us {100 * a[1] / tot:5.1f}%  read {a[2] / 1e9:6.3f} GB  written {a[3] / 1e9:6.3f} GB"
          f"  {(a[2] + a[3]) / a[1] / 1e3:6.0f} GB/s")
print(f"total {tot:.1f} us, {len(order)} launches, {sum(a[2] + a[3] for a in agg.values()) / 1e9:.3f} GB")
