#!/bin/bash
# Round 2: compute-sanitizer memcheck of small solves, ncu launch list of the 32-map share (256^2 x 32), final bench line.
mkdir -p gpurun_out; out=gpurun_out/r2_fifth.txt; : > $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?  $(grep -E 'ERROR SUMMARY|done' gpurun_out/sanitizer_memcheck.log | tr '\n' ' ')" >> $out
grep -E "Invalid|out of bounds|misaligned" gpurun_out/sanitizer_memcheck.log | head -10 >> $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_b32.csv python tools/quick_ms.py 256x32 --iters 3 > gpurun_out/ncu_b32.log 2>&1
echo "ncu b32 rc=$?" >> $out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_fifth.json 2> gpurun_out/bench_r2_fifth.err
echo "bench rc=$?" >> $out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_fifth_reference.json 2>> gpurun_out/bench_r2_fifth.err
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_fifth.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/bench_r2_fifth_reference.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1))
print('config equal to reference arm:', d['config'] == r['config'], '| reference', round(r['value'], 2), r['cpu_baseline']['kind'])
print('roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'] and d['roofline']['traffic_source'][:40])
PY
cat $out
