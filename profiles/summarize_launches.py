"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0][:64]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':64s} {'n':>4s} {'total us':>10s} {'share':>6s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:64s} {a[0]:4d} {a[1]:10.1f} {a[1] / tot * 100:5.1f}%")
print(f"{'TOTAL':64s} {sum(a[0] for a in agg.values()):4d} {tot:10.1f}")
