#!/bin/bash
# Round 2, final verification: parity suite, smoke(), bench line + reference arm, size table, launch lists.
mkdir -p gpurun_out; out=gpurun_out/r2_final.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_final.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_final.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_final.log | cut -c1-300 | head -30 >> $out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1
echo "smoke rc=$?  $(tail -2 gpurun_out/smoke_r2.log | cut -c1-250)" >> $out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_final3.json 2> gpurun_out/bench_r2_final3.err
echo "bench rc=$?" >> $out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final3_reference.json 2>> gpurun_out/bench_r2_final3.err
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_final3.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/bench_r2_final3_reference.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1), 'kernels', d['kernels_per_iteration'])
print('cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind']), 'gpu_eager', d['gpu_eager_baseline'] and d['gpu_eager_baseline'].get('value'), '| reference arm', round(r['value'], 2), 'config equal', d['config'] == r['config'])
print('others', json.dumps({k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in (d['other_configs'] or {}).items()}))
print('readme', d['readme_lens_ms_to_residual_1e-3'])
for x in [d['roofline']] + d['roofline_kernels']:
    print('  %-60s %8.1f us  %6.0f GB/s  %.3f' % (x['kernel'][:60], x['ms_per_launch'] * 1e3, x['achieved'], x['frac']))
print('stage unet', d['roofline_stage_unet']['stage_ms'], d['roofline_stage_unet']['frac'], 'spectral', d['roofline_stage_spectral']['stage_ms'], d['roofline_stage_spectral']['frac'])
PY
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 256x64 256x32 256x16 256x8 256x1 96x32 128x64 64x32 512x8 1024x1 --tag final >> $out 2>&1
for cfg in 96x32 1024x1; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_list_$cfg.csv python tools/quick_ms.py $cfg --iters 3 > gpurun_out/ncu_list_$cfg.log 2>&1
done
cat $out
