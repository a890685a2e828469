#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hi]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[hi + 1:]:
        d = dict(zip(H, r))
        a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"])
        tot += float(d["Metric Value"])
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.1f} us of device time (cold-cache, serialised)")
    for k, v in agg.items():
        print(f"{v[0]:5d} launches {v[1] / 1e3:10.1f} us total {v[1] / v[0] / 1e3:8.1f} us avg {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
