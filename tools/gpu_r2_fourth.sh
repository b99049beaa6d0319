#!/bin/bash
# Round 2, verification call: parity suite, smoke(), bench line, A/B of the overlapped chunked residual stage.
mkdir -p gpurun_out; out=gpurun_out/r2_fourth.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_fourth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_fourth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_fourth.log | cut -c1-300 | head -30 >> $out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1
echo "smoke rc=$?  $(tail -2 gpurun_out/smoke_r2.log | cut -c1-250)" >> $out
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 512x8 --tag base >> $out 2>&1
for ch in 16 32 64; do
    HELMNET_SPEC_CHUNK=$ch HELMNET_SPEC_OVERLAP=1 $q 256x256 256x128 --tag overlap_chunk$ch >> $out 2>&1
done
HELMNET_SPEC_CHUNK=2 HELMNET_SPEC_OVERLAP=1 $q 512x8 --tag overlap_chunk2 >> $out 2>&1
$q 256x256 --tag base_again >> $out 2>&1
HELMNET_SPEC_CHUNK=32 HELMNET_SPEC_OVERLAP=1 $q 256x256 --tag overlap_chunk32_again >> $out 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_fourth.json 2> gpurun_out/bench_r2_fourth.err
echo "bench rc=$?" >> $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2_fourth.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sustained', d['sustained'] and round(d['sustained']['value'], 1))
print('cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind']), 'gpu_eager', d['gpu_eager_baseline'] and d['gpu_eager_baseline'].get('value'))
print('stage unet', d['roofline_stage_unet']['stage_ms'], 'spectral', d['roofline_stage_spectral']['stage_ms'], d['roofline_stage_spectral']['frac'], 'roofline', d['roofline']['frac'])
PY
cat $out
