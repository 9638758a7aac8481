"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rN_launches_summary.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"# launch list summary ({sum(cnt.values())} launches, {total:.1f} ms of kernel time under ncu; shares, not absolutes)\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| `{k[:110]}` | {cnt[k]} | {v:.2f} | {100 * v / total:.1f}% |")
