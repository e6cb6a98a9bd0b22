"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/agg_launches.py file.csv"""
import csv
import re
import sys
from collections import defaultdict

tot, cnt = defaultdict(float), defaultdict(int)
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v  # -> us
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{t / 1e3:10.3f} ms {100 * t / total:6.2f}%  x{cnt[n]:<5d} {n}")
print(f"{total / 1e3:10.3f} ms total")
