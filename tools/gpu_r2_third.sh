#!/bin/bash
# Round 2, third GPU call: parity suite (sample groups, side branch, new fixtures), bench line with the eager-GPU baseline,
# A/B of sample groups, source-level ncu capture of the narrow fused DoubleConv kernels and decode[0].
mkdir -p gpurun_out; out=gpurun_out/r2_third.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_third.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_third.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_third.log | cut -c1-300 | head -40 >> $out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_third.json 2> gpurun_out/bench_r2_third.err
echo "bench rc=$?" >> $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_third.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1))
print('cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind']), 'gpu_eager', d['gpu_eager_baseline'])
print('others', json.dumps({k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in (d['other_configs'] or {}).items()}))
for r in [d['roofline']] + d['roofline_kernels']:
    print('  %-60s %8.1f us  %6.0f GB/s  %.3f' % (r['kernel'][:60], r['ms_per_launch'] * 1e3, r['achieved'], r['frac']))
print('stage unet', d['roofline_stage_unet']['stage_ms'], d['roofline_stage_unet']['frac'], 'spectral', d['roofline_stage_spectral']['stage_ms'], d['roofline_stage_spectral']['frac'])
PY
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 256x64 256x32 256x16 256x8 96x32 512x8 1024x1 --tag base >> $out 2>&1
HELMNET_GROUPS=2 $q 256x256 256x128 256x64 256x32 256x16 256x8 96x32 512x8 --tag groups2 >> $out 2>&1
HELMNET_GROUPS=4 $q 256x64 256x32 256x16 96x32 --tag groups4 >> $out 2>&1
HELMNET_GROUPS=2 HELMNET_SIDE_STATE=0 $q 256x64 256x32 --tag groups2_side0 >> $out 2>&1
HELMNET_GROUPS=2 HELMNET_PDL=0 $q 256x64 256x32 --tag groups2_pdl0 >> $out 2>&1
HELMNET_GROUPS=3 $q 256x32 --tag groups3 >> $out 2>&1
HELMNET_PDL=0 $q 256x32 256x16 --tag pdl0 >> $out 2>&1
$q 256x256 256x32 --tag base_again >> $out 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0"
# dconv launches of one iteration, in order: inc sig0 sta0 sig1 sta1 sig2 sta2 sig3 sta3 bot dec3 dec2 dec1 dec0 (14); capture
# sig2 .. dec0 of the fourth iteration with source
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dconv_tcf -s 47 -c 9 -o /tmp/r2_src2 $B > gpurun_out/ncu_src2.log 2>&1
echo "ncu src2 rc=$?" >> $out
ncu -i /tmp/r2_src2.ncu-rep --page source --csv > gpurun_out/r2_src2_source.csv 2>> gpurun_out/ncu_src2.log
ncu -i /tmp/r2_src2.ncu-rep --page raw --csv > gpurun_out/r2_src2_raw.csv 2>> gpurun_out/ncu_src2.log
ls -la gpurun_out | tail -8 >> $out
cat $out
