"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.

    python tools/summarize_launches.py gpurun_out/launches.csv STEPS > profiles/rNN_launches.md

STEPS = number of benchmark steps the profiled command executed (warm-up + timed + e2e), to print per-step figures.
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ns = v * 1e3 if unit.startswith("us") else (v * 1e6 if unit.startswith("ms") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    tot[name] += ns
    cnt[name] += 1
T = sum(tot.values())
print(f"| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v / T < 0.001:
        continue
    print(f"| `{k[:70]}` | {cnt[k] / steps:.1f} | {v / 1e6 / steps:.3f} | {100 * v / T:.1f}% |")
print(f"| **total** | {sum(cnt.values()) / steps:.1f} | {T / 1e6 / steps:.3f} | 100% |")
