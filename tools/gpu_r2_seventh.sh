#!/bin/bash
# Round 2: up-kernel strip heights from 2 rows (small batches), parity suite, final single-GPU bench line + launch list.
mkdir -p gpurun_out; out=gpurun_out/r2_seventh.txt; : > $out
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 256x64 256x32 256x16 256x8 256x1 96x32 128x64 64x32 512x8 1024x1 --tag final >> $out 2>&1
HELMNET_UP_ROWS=8 $q 256x32 256x8 256x1 96x32 --tag up_rows8 >> $out 2>&1
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_seventh.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_seventh.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_seventh.log | cut -c1-300 | head -30 >> $out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench rc=$?" >> $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1))
print('cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind']), 'gpu_eager', d['gpu_eager_baseline'] and d['gpu_eager_baseline'].get('value'))
print('others', json.dumps({k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in (d['other_configs'] or {}).items()}))
print('readme', d['readme_lens_ms_to_residual_1e-3'])
for r in [d['roofline']] + d['roofline_kernels']:
    print('  %-60s %8.1f us  %6.0f GB/s  %.3f' % (r['kernel'][:60], r['ms_per_launch'] * 1e3, r['achieved'], r['frac']))
print('stage unet', d['roofline_stage_unet']['stage_ms'], d['roofline_stage_unet']['frac'], 'spectral', d['roofline_stage_spectral']['stage_ms'], d['roofline_stage_spectral']['frac'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_b32.csv python tools/quick_ms.py 256x32 --iters 3 > gpurun_out/ncu_b32.log 2>&1
echo "ncu b32 rc=$?" >> $out
cat $out
