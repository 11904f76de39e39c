#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step).
usage: python tests/ubench/summarize_launches.py gpurun_out/<tag>_launches.csv > profiles/<tag>_ncu_launches_summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
agg = collections.OrderedDict()
for r in rows:
    k = re.sub(r'\(.*', '', r['Kernel Name'])[:72]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r['Metric Value'].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none; bench.py --steps 1 --warmup 3 (32 scenes; profiler range = the timed step)")
print("# cold-cache serialised launches: compare SHARES.  total %.3f ms over %d launches" % (tot / 1e6, len(rows)))
print("%-74s %6s %12s %8s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-74s %6d %12.1f %8.1f %6.1f%%" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
