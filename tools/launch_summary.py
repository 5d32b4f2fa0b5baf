#!/usr/bin/env python
"""Per-kernel launch count / average duration from an ncu `--metrics gpu__time_duration.sum --csv` log.
usage: python tools/launch_summary.py gpurun_out/launches_X.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, st = r, i
        break
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[st + 1:]:
    if len(r) > vi:
        a = agg.setdefault(r[ki].split('(')[0], [0, 0.0])
        a[0] += 1; a[1] += float(r[vi].replace(',', ''))
for k, a in agg.items():
    print('%-40s %4d launches  avg %9.1f us  total %9.1f us' % (k, a[0], a[1] / a[0] / 1e3, a[1] / 1e3))
