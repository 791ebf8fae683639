"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.  Usage: launch_summary.py file.csv [top]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for x in rows:
    name = re.sub(r'\(.*', '', x['Kernel Name'])
    v = float(x['Metric Value'].replace(',', ''))
    v = v / 1e3 if x['Metric Unit'] == 'ns' else v * 1e3 if x['Metric Unit'] == 'ms' else v
    a = agg[name]
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(v[1] for v in agg.values())
print('launches %d, total %.1f us' % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%-52s n=%5d total %9.1f us  avg %7.2f  max %7.2f us  %5.1f%%' % (k[:52], v[0], v[1], v[1] / v[0], v[2], 100 * v[1] / tot))
