"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3 if unit in ("ms", "msecond") else val
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, u in rows:
    agg[n][0] += 1; agg[n][1] += u
print(f"# {path}: {len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:8d} {u:12.1f} {u/c:10.2f} {100*u/tot:6.1f}%")
