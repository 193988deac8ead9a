"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ai, si, ni, xi = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
items = []
for k, r in enumerate(rows[2:]):
    try:
        items.append((float(r[ni]), float(r[xi]), k, r[si].strip()))
    except Exception:
        pass
tot = sum(i[0] for i in items)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("total samples", tot, "instructions", len(items))
for v, x, k, s in sorted(items, reverse=True)[:n]:
    print(f"{v:8.0f} {v / tot * 100:5.1f}%  exec={x:9.0f}  #{k:4d}  {s[:100]}")
