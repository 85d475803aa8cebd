"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    if row["Metric Unit"] == "us":
        v *= 1e3
    elif row["Metric Unit"] == "ms":
        v *= 1e6
    agg.setdefault(row["Kernel Name"].split("(")[0][-50:], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:52s} n={len(v):4d} mean={sum(v)/len(v)/1e3:10.2f} us  share={100*sum(v)/tot:5.1f}%")
