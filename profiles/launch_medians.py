#!/usr/bin/env python3
"""Median duration and DRAM bytes per kernel of an ncu launch list (gpu__time_duration.sum, dram__bytes_*.sum).
usage: launch_medians.py <launches.csv>"""
import csv
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"name": r[4].split("(")[0]})
        d[r[12]] = float(r[14].replace(",", ""))
    agg = defaultdict(list)
    for i in sorted(launches):
        l = launches[i]
        agg[l["name"]].append((l["gpu__time_duration.sum"], l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)))
    print(f"{'kernel':44s} {'n':>5s} {'median us':>10s} {'dram MB':>9s} {'dram GB/s':>10s}")
    for k, v in agg.items():
        v = [x for x in v if x[0] > 3000]  # drop launches that return at once (e.g. the fallback histogram pass)
        if not v:
            continue
        t, b = sorted(v)[len(v) // 2]
        print(f"{k:44s} {len(v):5d} {t / 1e3:10.1f} {b / 1e6:9.1f} {b / t:10.1f}")


if __name__ == "__main__":
    main()
