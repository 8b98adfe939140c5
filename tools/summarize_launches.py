"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (development aid)."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^(void )?vmx::", "", name)
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    t = tot[name]
    t[0] += 1
    t[1] += ms
    t[2] = max(t[2], ms)
total = sum(t[1] for t in tot.values())
print("# %d launches, %.1f ms total" % (sum(t[0] for t in tot.values()), total))
print("%-34s %8s %12s %8s %12s" % ("kernel", "launches", "total_ms", "share", "max_ms"))
for name, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    if t[1] / total < 0.0002:
        continue
    print("%-34s %8d %12.3f %7.1f%% %12.3f" % (name[:34], t[0], t[1], 100 * t[1] / total, t[2]))
