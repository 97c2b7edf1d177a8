"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"<unnamed>::|void ", "", r["Kernel Name"]).split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r["Metric Unit"], v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"launches {sum(v[0] for v in agg.values())}, summed kernel time {tot:.0f} us (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.0f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg {v[1] / v[0]:7.1f} us  {k[:120]}")
