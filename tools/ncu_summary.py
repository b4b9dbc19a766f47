"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / total / mean."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit == "ns": v /= 1e3
    elif unit == "ms": v *= 1e3
    elif unit == "s": v *= 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'count':>7s} {'total_us':>10s} {'mean_us':>8s} {'share':>6s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {c:7d} {t:10.1f} {t/c:8.2f} {100*t/tot:5.1f}%")
print(f"{'TOTAL':70s} {sum(a[0] for a in agg.values()):7d} {tot:10.1f}")
