"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name (and grid for GEMMs)."""
import csv
import collections
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = r["Kernel Name"].split("(")[0]
    rows.append((name, r.get("Grid Size", ""), v))
tot = sum(v for _, _, v in rows)
by = collections.defaultdict(lambda: [0, 0.0])
for n, g, v in rows:
    by[n][0] += 1
    by[n][1] += v
print(f"total {tot/1e3:.2f} ms over {len(rows)} launches")
for n, (c, v) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/tot*100:6.2f}%  {v/1e3:9.3f} ms  n={c:5d}  avg={v/c:9.1f} us  {n}")
if len(sys.argv) > 2:
    byg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, v in rows:
        if sys.argv[2] in n:
            byg[g][0] += 1
            byg[g][1] += v
    for g, (c, v) in sorted(byg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"   grid {g:>18s}  n={c:4d}  avg={v/c:9.1f} us  tot={v/1e3:8.3f} ms")
