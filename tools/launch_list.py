#!/usr/bin/env python
"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list compactly: id, kernel, grid, block, microseconds."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = 0.0
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        us = float(x["Metric Value"].replace(",", "")) / (1e3 if x["Metric Unit"] == "ns" else 1.0)
        tot += us
        print(f'{x["ID"]:>4} {x["Kernel Name"][:48]:48} {x["Grid Size"]:>14} {x["Block Size"]:>14} {us:9.2f}')
print(f"total {tot:.2f} us")
