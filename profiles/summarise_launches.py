"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel, launches, total ms, share.

    python profiles/summarise_launches.py gpurun_out/launches.csv > profiles/rX_launch_summary.csv
"""
import collections
import csv
import re
import sys

tot = collections.Counter()
cnt = collections.Counter()
with open(sys.argv[1], newline="") as fh:
    rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*\)$", "", name).replace("b200seed::", "").replace("b200seed_rx::", "rx::")
    tot[name] += float(r[iv].replace(",", ""))
    cnt[name] += 1
total = sum(tot.values())
print("kernel,launches,total_ms,share")
for k, v in tot.most_common():
    print('"%s",%d,%.3f,%.2f%%' % (k, cnt[k], v / 1e6, 100 * v / total))
