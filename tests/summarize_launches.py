"""Developer tool: summarise an ncu launch list (gpu__time_duration.sum csv) by kernel and by (kernel, grid)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
by_k, by_kg = collections.defaultdict(list), collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", "")) / 1e3
    except ValueError:
        continue
    n = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")
    by_k[n].append(v)
    by_kg[(n, r[gi])].append(v)
tot = sum(sum(v) for v in by_k.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in by_k.values())} launches")
for k, v in sorted(by_k.items(), key=lambda x: -sum(x[1])):
    print(f"{sum(v):9.1f} us {100 * sum(v) / tot:5.1f}% n={len(v):4d} {k}")
if len(sys.argv) > 2:
    print("--- by (kernel, grid)")
    for k, v in sorted(by_kg.items(), key=lambda x: -sum(x[1]))[:int(sys.argv[2])]:
        print(f"{sum(v):9.1f} us n={len(v):4d} avg={sum(v) / len(v):7.1f} {k[0]} {k[1]}")
