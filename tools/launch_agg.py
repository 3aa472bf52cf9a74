"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: python tools/launch_agg.py launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])[:100]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1]) / 1e6
for k, v in agg.items():
    print(f"{v[1]:9.3f} ms {v[0]:5d}  {k}")
