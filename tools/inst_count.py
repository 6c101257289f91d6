"""Sums an ncu --csv metric log per kernel name: python tools/inst_count.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ix = {k: j for j, k in enumerate(h)}
tot = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows[start:]:
    if len(r) < len(h):
        continue
    name = r[ix["Kernel Name"]].split("(")[0] + " x" + r[ix["Block Size"]].strip("()").split(",")[0]
    tot[name][r[ix["Metric Name"]]] += float(r[ix["Metric Value"]].replace(",", ""))
for name, m in sorted(tot.items(), key=lambda kv: -sum(kv[1].values())):
    print(name, {k: f"{v:.4g}" for k, v in m.items()})
