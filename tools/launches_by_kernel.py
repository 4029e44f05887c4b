"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python tools/launches_by_kernel.py launches.csv > profiles/launches_<round>_by_kernel.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000.0 if row["Metric Unit"] == "ns" else (v * 1000.0 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(row["Kernel Name"][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(t for _, t in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "mean_us", "total_us", "share"])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, "%.1f" % (t / n), "%.1f" % t, "%.3f" % (t / total)])
