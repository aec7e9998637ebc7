#!/usr/bin/env python3
"""Per-kernel totals and shares of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file list.csv ...).
usage: launch_summary.py list.csv commits"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, commits = sys.argv[1], int(sys.argv[2])
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*$", "", r[name_i])[:62]
        v = float(r[val_i].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[unit_i], 1e-6)  # -> ms
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-62s launches %4d  total %9.3f ms  per commit %8.3f ms  share %5.1f%%" % (name, n, t, t / commits, 100 * t / total))
    print("total %.3f ms over %d commits = %.1f ms per commit" % (total, commits, total / commits))


if __name__ == "__main__":
    main()
