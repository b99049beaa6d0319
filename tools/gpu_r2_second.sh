#!/bin/bash
# Round 2, second GPU call: parity suite with the new defaults, bench line, reference arm, A/B of deeper narrow operand rings
# (tools/libhelmnet_alt.so), ncu launch list + full capture (exported to CSV on the box: the .ncu-rep is too big to travel back).
mkdir -p gpurun_out; out=gpurun_out/r2_second.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_second.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_second.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_second.log | cut -c1-300 | head -40 >> $out
if grep -q "^FAILED" gpurun_out/tests_r2_second.log; then
    HELMNET_TCD_MIN_RES=64 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_second_tcd64.log 2>&1
    echo "tests (TCD_MIN_RES=64) rc=$?  $(tail -1 gpurun_out/tests_r2_second_tcd64.log)" | tee -a $out
    grep -E "^FAILED" gpurun_out/tests_r2_second_tcd64.log | head -20 >> $out
fi
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_second.json 2> gpurun_out/bench_r2_second.err
echo "bench rc=$?" >> $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_second.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1))
print('cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind']), 'gpu_eager', d['gpu_eager_baseline'])
print('others', json.dumps(d['other_configs']))
print('ttr', d['ms_to_residual_1e-3'], d['readme_lens_ms_to_residual_1e-3'])
for r in [d['roofline']] + d['roofline_kernels']:
    print('  %-60s %8.1f us  %6.0f GB/s  %.3f' % (r['kernel'][:60], r['ms_per_launch'] * 1e3, r['achieved'], r['frac']))
print('stage unet', d['roofline_stage_unet']['stage_ms'], d['roofline_stage_unet']['frac'], 'spectral', d['roofline_stage_spectral']['stage_ms'], d['roofline_stage_spectral']['frac'])
print('clocks', d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2_second.err
echo "reference arm rc=$? $(cut -c1-400 gpurun_out/bench_r2_reference.json)" >> $out
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x32 64x256 64x32 32x256 128x256 96x32 --tag base >> $out 2>&1
HELMNET_SM100_LIB=$PWD/tools/libhelmnet_alt.so $q 256x256 256x32 64x256 64x32 32x256 128x256 96x32 --tag alt_rings >> $out 2>&1
HELMNET_TCD_MIN_RES=64 $q 256x256 256x32 --tag tcd64 >> $out 2>&1
HELMNET_SIDE_STATE=1 $q 256x256 256x128 --tag side1 >> $out 2>&1
HELMNET_SIDE_STATE=0 $q 256x256 256x64 256x32 --tag side0 >> $out 2>&1
$q 256x256 256x32 --tag base_again >> $out 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> $out
timeout 900 ncu --set full --clock-control none -s 58 -c 30 -o /tmp/r2_full $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> $out
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>> gpurun_out/ncu_full.log
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'dconv_tcf_kernel<2, 0, 0>|dconv_tcf_kernel<3, 0, 0>|spectral_cols256' -s 6 -c 3 -o /tmp/r2_src $B > gpurun_out/ncu_src.log 2>&1
echo "ncu src rc=$?" >> $out
ncu -i /tmp/r2_src.ncu-rep --page source --csv > gpurun_out/r2_src_source.csv 2>> gpurun_out/ncu_src.log
ncu -i /tmp/r2_src.ncu-rep --page raw --csv > gpurun_out/r2_src_raw.csv 2>> gpurun_out/ncu_src.log
ls -la gpurun_out >> $out
cat $out
