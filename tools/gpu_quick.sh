#!/bin/bash
# One short gpurun call: GPU parity suite + the per-kernel timing table of the headline config + two smaller sizes.
mkdir -p gpurun_out; out=gpurun_out/quick.txt; : > $out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/tests_quick.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_quick.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_quick.log | head -20 >> $out
tools/bench_kernels.sh headline >> $out 2>&1
for nb in "128 64" "128 256" "192 32"; do
    set -- $nb
    timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
        python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'])
for r in d['roofline_kernels']:
    if 'down' in r['kernel']: print('   ', r['kernel'], round(r['ms_per_launch']*1e3,1), 'us', round(r['frac'],3))
" >> $out 2>&1
done
cat $out
