"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
python tools/launch_summary.py gpurun_out/x.csv [skip_first_n] > profiles/x_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v * scale))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    agg[k][0] += 1
    agg[k][1] += us
total = sum(v[1] for v in agg.values())
print(f"# {path}: {len(rows)} launches, {total / 1e3:.2f} ms of kernel time (ncu: serialised, cold caches)")
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:9d} {us / 1e3:10.3f} {100 * us / total:6.1f}% {us / n:9.1f}")
