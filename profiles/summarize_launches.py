#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path, first=None, last=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if first is not None:
        rows = rows[int(first):int(last)]
    agg = collections.OrderedDict()
    for row in rows:
        k = row["Kernel Name"].split("(")[0][:48]
        agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("%d launches, total %.3f ms (cold-cache, serialised: compare shares)" % (len(rows), tot / 1e6))
    for k, v in agg.items():
        print("%-50s n=%3d total=%9.3f ms avg=%8.3f ms share=%5.1f%%" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6,
                                                                          100 * sum(v) / tot))


if __name__ == "__main__":
    main(*sys.argv[1:])
