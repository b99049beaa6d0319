#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one block of key metrics per captured kernel launch."""
import csv
import sys

SEL = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
       'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
       'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def find(k):
        for h in hdr:
            if h.endswith(k):
                return h

    for r in rows[2:]:
        print('=====', r[idx['Kernel Name']][:70], r[idx['Grid Size']])
        for k in SEL:
            h = find(k)
            if h:
                print(f"   {k:85s} {r[idx[h]]} {units[idx[h]]}")
        for h in hdr:
            if 'average_warps_issue_stalled' in h and r[idx[h]] not in ('', '0'):
                try:
                    if float(r[idx[h]]) < 0.05:
                        continue
                except ValueError:
                    pass
                print(f"      stall {h.split('stalled_')[1][:40]:42s} {r[idx[h]]}")


if __name__ == '__main__':
    main(sys.argv[1])
