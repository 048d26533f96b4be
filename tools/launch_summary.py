"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
names = rows[hdr]
ki, vi, ui = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    k = r[ki].split("(")[0][:70]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg.setdefault(k, []).append(v)
for k, v in agg.items():
    if k.startswith("void at::"):
        continue
    print(f"{k:72s} n={len(v):4d} mean={sum(v) / len(v):8.2f}us total={sum(v) / 1000:8.3f}ms")
