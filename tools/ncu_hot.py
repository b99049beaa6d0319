#!/usr/bin/env python
"""Hot SASS instructions (by warp-stall samples) per kernel from an `ncu --page source --csv` export."""
import csv
import sys


def main(path, pattern, thresh=0.01):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == 'Kernel Name':
            name = r[1]
            hdr = rows[i + 1]
            j = i + 2
            data = []
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                if len(rows[j]) == len(hdr):
                    data.append(rows[j])
                j += 1
            if pattern in name:
                si, ci = hdr.index('Source'), hdr.index('# Samples')
                ii = hdr.index('Instructions Executed')
                tot = sum(float(d[ci]) for d in data)
                print('=====', name, 'samples', tot, 'instructions', len(data))
                for k, d in enumerate(data):
                    v = float(d[ci])
                    if v / tot > thresh:
                        print(f"{k:5d} {100 * v / tot:6.2f}%  exec={d[ii]:>9s} {d[si][:100]}")
                return
            i = j
        else:
            i += 1


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.01)
