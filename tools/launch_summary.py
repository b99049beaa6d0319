#!/usr/bin/env python
"""Per-kernel average duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path, flt=None):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        key = (d['Kernel Name'][:64], d['Grid Size'])
        agg.setdefault(key, []).append(float(d['Metric Value'].replace(',', '')))
    tot = 0.0
    for k, v in agg.items():
        if flt and flt not in k[1] and flt not in k[0]:
            continue
        print(f"{k[0]:64s} grid={k[1]:14s} n={len(v):3d} avg_us={sum(v) / len(v) / 1e3:9.1f}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
