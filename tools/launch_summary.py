#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/r01_launches.txt"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    agg = collections.OrderedDict()
    for row in rd:
        d = dict(zip(hdr, row))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
        key = (name, d["Grid Size"] if False else "")
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print("# source: %s  (per-launch times are cold-cache and serialised: compare SHARES)" % path)
    print("%-44s %7s %12s %8s %12s %12s" % ("kernel", "count", "total ms", "share", "avg ms", "max ms"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %7d %12.3f %7.1f%% %12.4f %12.4f" % (k[:44], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2]))
    print("%-44s %7d %12.3f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
