"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>9s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {t:12.1f} {100*t/tot:6.1f}% {t/n:9.1f}")
print(f"{'TOTAL':60s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")
