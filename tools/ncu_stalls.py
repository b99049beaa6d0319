#!/usr/bin/env python
"""Per-kernel warp-stall summary and hottest SASS instructions from an `ncu --page source --csv` export
(`ncu --set full --import-source on ...` capture).

    python tools/ncu_stalls.py gpurun_out/r2_src2_source.csv [name-substring] [top N] > profiles/r2_ncu_dconv_stalls.txt
"""
import csv
import sys


def blocks_of(path):
    rows = list(csv.reader(open(path)))
    i, out = 0, []
    while i < len(rows):
        r = rows[i]
        if r and r[0] == 'Kernel Name':
            name, hdr, j, data = r[1], rows[i + 1], i + 2, []
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                if len(rows[j]) == len(hdr):
                    data.append(rows[j])
                j += 1
            out.append((name, hdr, data))
            i = j
        else:
            i += 1
    return out


def main():
    path = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ''
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    seen = set()
    print(f"# {path}: warp-stall sampling per kernel (ncu --set full --import-source on --clock-control none), hottest SASS instructions")
    print("# long_sb = long scoreboard (global / L2 / mbarrier try_wait), short_sb = shared memory / TMEM, wait = fixed-latency dependency")
    for name, hdr, data in blocks_of(path):
        if pat not in name or name in seen:
            continue
        seen.add(name)
        H = {h: k for k, h in enumerate(hdr)}
        st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(float(d[H['# Samples']]) for d in data) or 1.0
        agg = {s: sum(float(d[H[s]] or 0) for d in data) for s in st}
        print(f"\n== {name[:110]}\n   {int(tot)} samples, {len(data)} instructions")
        print("   " + "  ".join(f"{s[6:]} {100 * v / tot:.0f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        for k, d in sorted(enumerate(data), key=lambda x: -float(x[1][H['# Samples']]))[:top_n]:
            v = float(d[H['# Samples']])
            top = sorted(((s, float(d[H[s]] or 0)) for s in st), key=lambda x: -x[1])[:2]
            print(f"   {k:5d} {100 * v / tot:5.2f}%  exec {d[H['Instructions Executed']]:>8s}  {d[H['Source']][:72]:72s} {top[0][0][6:]} {top[0][1]:.0f}, {top[1][0][6:]} {top[1][1]:.0f}")


if __name__ == '__main__':
    main()
