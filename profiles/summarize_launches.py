"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
rows = [r for r in rows if r[0] >= skip]
agg = defaultdict(lambda: [0, 0.0])
for _, name, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = short if len(short) < 90 else short[:87] + "..."
    agg[short][0] += 1
    agg[short][1] += ns
total = sum(v[1] for v in agg.values())
mine = sum(v[1] for k, v in agg.items() if k.startswith("chap::"))
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / total))
print("\ntotal %.3f ms over %d launches; kernels of this repository (chap::*): %.1f%% of device time, %d launches"
      % (total / 1e6, len(rows), 100 * mine / total, sum(v[0] for k, v in agg.items() if k.startswith("chap::"))))
