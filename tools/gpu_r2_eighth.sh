#!/bin/bash
# Round 2: 4-line tiles of the N = 256 residual kernels for small solves, iteration-slot advance folded into the first kernel,
# PDL / side branch up to 8 Mi points: A/B, parity suite, bench line.
mkdir -p gpurun_out; out=gpurun_out/r2_eighth.txt; : > $out
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 256x64 256x32 256x16 256x8 256x1 96x32 --tag auto >> $out 2>&1
HELMNET_SPEC_LINES=8 $q 256x128 256x64 256x32 256x16 256x8 256x1 --tag lines8 >> $out 2>&1
HELMNET_SPEC_LINES=4 $q 256x256 256x128 256x64 --tag lines4 >> $out 2>&1
$q 256x256 256x32 --tag auto_again >> $out 2>&1
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_eighth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_eighth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_eighth.log | cut -c1-300 | head -30 >> $out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err
echo "bench rc=$?" >> $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_final2.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1), 'kernels', d['kernels_per_iteration'], 'launches', d['gpu_launches'])
print('others', json.dumps({k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in (d['other_configs'] or {}).items()}))
print('readme', d['readme_lens_ms_to_residual_1e-3'] and d['readme_lens_ms_to_residual_1e-3']['ms'])
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_b32.csv python tools/quick_ms.py 256x32 --iters 3 > gpurun_out/ncu_b32.log 2>&1
echo "ncu b32 rc=$?" >> $out
cat $out
