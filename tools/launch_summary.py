#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total, median, share.

    python tools/launch_summary.py profiles/launches_r02_ekf20k.csv
"""
import collections
import csv
import re
import statistics
import sys


def main():
    lines = open(sys.argv[1]).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines[start:]):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(row["Metric Unit"], 1.0)
        agg[re.sub(r"\(.*", "", row["Kernel Name"])].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':60s} {'count':>5s} {'total ms':>10s} {'median us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        print(f"{k[:60]:60s} {len(v):5d} {sum(v) / 1e3:10.3f} {statistics.median(v):10.2f} {sum(v) / tot * 100:5.1f}%")


if __name__ == "__main__":
    main()
