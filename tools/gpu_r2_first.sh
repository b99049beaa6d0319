#!/bin/bash
# Round 2, first GPU call: parity suite, default bench line, environment A/Bs, ncu launch list + full capture with source.
mkdir -p gpurun_out; out=gpurun_out/r2_first.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_first.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_first.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_first.log | head -30 >> $out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err
echo "bench rc=$?" >> $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_first.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1))
print('cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind']), 'gpu_eager', d['gpu_eager_baseline'])
print('others', json.dumps(d['other_configs']))
for r in [d['roofline']] + d['roofline_kernels']:
    print('  %-60s %8.1f us  %6.0f GB/s  %.3f' % (r['kernel'][:60], r['ms_per_launch'] * 1e3, r['achieved'], r['frac']))
print('stage unet', d['roofline_stage_unet']['stage_ms'], d['roofline_stage_unet']['frac'], 'spectral', d['roofline_stage_spectral']['stage_ms'], d['roofline_stage_spectral']['frac'])
print('clocks', d['clocks'])
PY
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 256x64 256x32 256x16 256x1 96x32 --tag base >> $out 2>&1
HELMNET_SPEC_CHUNK=32 $q 256x256 --tag spec_chunk32 >> $out 2>&1
HELMNET_SPEC_CHUNK=64 $q 256x256 --tag spec_chunk64 >> $out 2>&1
HELMNET_TCD_MIN_RES=32 $q 256x256 256x32 --tag tcd32 >> $out 2>&1
HELMNET_TCD_MIN_RES=16 $q 256x256 256x32 --tag tcd16 >> $out 2>&1
HELMNET_PDL=0 $q 256x32 256x64 --tag pdl0 >> $out 2>&1
HELMNET_PDL=1 $q 256x32 256x64 256x256 --tag pdl1 >> $out 2>&1
HELMNET_PDL=2 $q 256x64 256x256 --tag pdl2 >> $out 2>&1
HELMNET_DCONV_MIN_ROWS=8 $q 256x32 --tag minrows8 >> $out 2>&1
HELMNET_FUSE_BOTTOM=0 $q 256x256 256x32 256x1 --tag nofusebottom >> $out 2>&1
HELMNET_SRC_SKIP=0 $q 256x256 256x32 --tag nosrcskip >> $out 2>&1
HELMNET_SIDE_STATE=1 $q 256x256 256x64 256x32 256x16 256x1 96x32 --tag side_state >> $out 2>&1
HELMNET_SIDE_STATE=1 HELMNET_PDL=1 $q 256x32 --tag side_state_pdl1 >> $out 2>&1
# ncu: launch list of the bench command (serialised, cold cache), then a full capture with source of one iteration
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> $out
timeout 900 ncu --set full --clock-control none -s 60 -c 32 -o gpurun_out/r2_full $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?  $(ls -la gpurun_out/r2_full.ncu-rep 2>/dev/null)" >> $out
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'dconv_tcf_kernel<2, 0, 0>|dconv_tcf_kernel<3, 0, 0>|spectral_cols256|dconv_tcf_kernel<3, 2, 1>' -s 8 -c 4 -o gpurun_out/r2_src $B > gpurun_out/ncu_src.log 2>&1
echo "ncu src rc=$?  $(ls -la gpurun_out/r2_src.ncu-rep 2>/dev/null)" >> $out
cat $out
