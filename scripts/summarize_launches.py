#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per kernel: launches, total and average device time, share of the total."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("%-64s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_ms", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-64s %6d %12.3f %10.4f %6.1f%%" % (k[:64], a[0], a[1], a[1] / a[0], 100 * a[1] / total))
    print("%-64s %6s %12.3f" % ("TOTAL", "", total))


if __name__ == "__main__":
    main(sys.argv[1])
