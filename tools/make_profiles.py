#!/usr/bin/env python
"""Turn the raw ncu outputs of one gpurun capture into the committed summaries under profiles/.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e2.csv \\
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --residual-iters 0
    ncu --set full --clock-control none --import-source on -k regex:"dconv_tcf|spectral|down_tcr|up_tcr" -s 14 -c 14 \\
        -o gpurun_out/prof_e2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --residual-iters 0
    ncu -i gpurun_out/prof_e2.ncu-rep --page raw --csv > gpurun_out/prof_e2_raw.csv
    python tools/make_profiles.py gpurun_out/launches_e2.csv gpurun_out/prof_e2_raw.csv <bench ms per iteration> [r2]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS = 148
PREFIX = sys.argv[4] if len(sys.argv) > 4 else 'r1'       # round prefix of the files written under profiles/


def launches(path, bench_us):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    # the last complete iteration: from one reset_amax_kernel (the first kernel of an iteration) to the launch before the next one
    idx = [i for i, d in enumerate(data) if 'reset_amax' in d['Kernel Name']]
    a, b = idx[-2] - 1, idx[-1] - 1
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3` (256^2 x 256, engine 2), one B200.",
           "# One solver iteration = the launches from one reset_amax_kernel to the next (replayed as one CUDA graph in production).",
           "# Per-launch times are cold-cache and serialised by ncu: the kernel's SHARE of the iteration is what compares with bench.py.",
           f"# {'kernel':62s} {'grid':14s} {'block':12s} {'us':>8s} {'share':>6s}"]
    tot = sum(float(d['Metric Value'].replace(',', '')) / 1e3 for d in data[a + 1:b + 1])
    for d in data[a + 1:b + 1]:
        t = float(d['Metric Value'].replace(',', '')) / 1e3
        out.append(f"{d['Kernel Name'][:64]:64s} {d['Grid Size']:14s} {d['Block Size']:12s} {t:8.1f} {100 * t / tot:5.1f}%")
    out.append(f"# total {tot:.1f} us per iteration under ncu ({b - a} launches); bench.py (graph replay, warm): {bench_us:.0f} us")
    open(os.path.join(ROOT, 'profiles', f'{PREFIX}_launches_engine2.txt'), 'w').write("\n".join(out) + "\n")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]

    def col(k):
        hs = [h for h in hdr if h.endswith(k)]
        return hdr.index(hs[0]) if hs else None

    want = [('us', 'gpu__time_duration.sum'), ('cyc', 'sm__cycles_elapsed.max'), ('tc_wf', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum'),
            ('lsu_wf', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'), ('ld_confl', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum'),
            ('st_confl', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum'), ('dram_rd', 'dram__bytes_read.sum'),
            ('dram_wr', 'dram__bytes_write.sum'), ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            ('tc_pipe%', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed'),
            ('tensor%', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
            ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), ('regs', 'launch__registers_per_thread'),
            ('warps%', 'sm__warps_active.avg.pct_of_peak_sustained_active'), ('utcmma', 'smsp__sass_inst_executed_op_utcmma.sum')]
    out = ["# ncu --set full --clock-control none, one launch of each level-0/1 kernel of the engine-2 iteration, 256^2 x 256, one B200.",
           "# smem model: the SM's shared memory moves one 128-byte wavefront per cycle for ALL clients; wavefronts per SM =",
           "#   (tcgen05 operand reads + LSU wavefronts without bank-conflict replays + TMA fill bytes/128) ~ elapsed cycles for the fused",
           "#   DoubleConv kernels, i.e. they run at the shared-memory bandwidth roofline (see DESIGN.md 4.2).", ""]
    seen, traffic = set(), {}
    ids = {'dconv_tcf_kernel<3, 2, 1>': 9, 'dconv_tcf_kernel<2, 2, 0>': 7, 'dconv_tcf_kernel<2, 2, 2>': 8, 'dconv_tcf_kernel<0, 2, 0>': 6,
           'spectral_rows256': 4, 'spectral_cols256': 5}
    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    best = {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        key = (name, r[hdr.index('Grid Size')])
        rd = float(r[col('dram__bytes_read.sum')]) * mult[units[col('dram__bytes_read.sum')]]
        wr = float(r[col('dram__bytes_write.sum')]) * mult[units[col('dram__bytes_write.sum')]]
        for k, i in ids.items():
            if k in name and i not in traffic:
                traffic[i] = {"kernel": name[:60], "dram_bytes_read": rd, "dram_bytes_write": wr}
        for k, i in (('down_tcr', 2), ('up_tcr', 3)):
            if k in name and (i not in best or rd + wr > best[i][0]):
                best[i] = (rd + wr, {"kernel": name[:60], "dram_bytes_read": rd, "dram_bytes_write": wr})
        if key in seen:
            continue
        seen.add(key)
        out.append(f"== {name[:70]} grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        vals = {}
        for k, m in want:
            c = col(m)
            if c is None:
                continue
            vals[k] = r[c]
            out.append(f"   {m:80s} {r[c]} {units[c]}")
        try:
            cyc = float(vals['cyc'])
            tc = float(vals['tc_wf']) / SMS
            lsu = (float(vals['lsu_wf']) - float(vals['ld_confl']) - float(vals['st_confl'])) / SMS
            out.append(f"   -> per SM: tcgen05 operand wavefronts {tc:,.0f} + LSU wavefronts (no replays) {lsu:,.0f} = {tc + lsu:,.0f} of "
                       f"{cyc:,.0f} elapsed cycles ({100 * (tc + lsu) / cyc:.0f} %, TMA fills not counted)")
            n_mma = float(vals.get('utcmma', '0') or 0) / SMS
            if n_mma > 0:
                out.append(f"   -> {n_mma:,.0f} tcgen05.mma per SM, {cyc / n_mma:.1f} cycles per MMA (44.5 in isolation, tools/tc_bench.cu)")
        except (KeyError, ValueError):
            pass
    open(os.path.join(ROOT, 'profiles', f'{PREFIX}_ncu_full_engine2.txt'), 'w').write("\n".join(out) + "\n")
    for i, (_, v) in best.items():
        traffic[i] = v
    doc = {"config": {"n": 256, "batch_per_gpu": 256},
           "source": f"ncu --set full --clock-control none, profiles/{PREFIX}_ncu_full_engine2.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                     "in-iteration launches)",
           "kernels": {str(k): v for k, v in sorted(traffic.items())}}
    json.dump(doc, open(os.path.join(ROOT, 'profiles', f'{PREFIX}_dram_traffic.json'), 'w'), indent=1)


if __name__ == '__main__':
    launches(sys.argv[1], float(sys.argv[3]) * 1e3)
    full(sys.argv[2])
