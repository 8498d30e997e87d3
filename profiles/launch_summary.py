#!/usr/bin/env python
"""Per-kernel summary (launch count, mean / max duration) of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[ki][:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k:90s} n={len(v):3d} mean={sum(v) / len(v) / 1e6:9.3f} ms  max={max(v) / 1e6:9.3f} ms  share={100 * sum(v) / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
