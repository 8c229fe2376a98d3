"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the ordered list."""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path, errors="replace")))
hdr = None
order, agg = [], collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    val = float(d["Metric Value"].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(d["Metric Unit"], 1.0)
    name = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")[:70]
    grid = d.get("Grid Size", "")
    order.append((name, val * scale, grid))
    agg[name][0] += 1
    agg[name][1] += val * scale
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(order)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100*v[1]/tot:5.1f}%  x{v[0]:4d}  avg {v[1]/v[0]:9.1f} us  {k}")
if len(sys.argv) > 2:
    print("--- ordered ---")
    for name, us, grid in order:
        print(f"{us:10.1f} us  grid {grid:>14}  {name}")
