"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: agg_launches.py file.csv"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    rows.append(row)
agg = collections.OrderedDict()
tot = 0.0
for row in rows:
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:78]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} us {n:4d}x {t/n:8.1f} us/launch {100*t/tot:5.1f}%  {k}")
print(f"total {tot:.1f} us over {len(rows)} launches")
