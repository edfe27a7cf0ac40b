#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: tools/agg_launches.py launches.csv [first-kernel-substring last-kernel-substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
seq = [(r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-44:], float(r[vi].replace(",", "")) / 1000)
       for r in rows[hdr + 1:] if len(r) > vi]
if len(sys.argv) > 3:
    last = max(i for i, (k, v) in enumerate(seq) if sys.argv[3] in k)
    first = max(i for i, (k, v) in enumerate(seq[:last]) if sys.argv[2] in k)
    seq = seq[first:last + 1]
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
print("launches: %d, sum %.1f us" % (len(seq), sum(v for k, v in seq)))
for k, (n, v) in agg.items():
    print("  %-44s x%-4d %9.1f us" % (k, n, v))
