#!/bin/bash
# One gpurun call: parity suite with the any-width fused DoubleConv kernels, then A/B against the four exact widths only.
mkdir -p gpurun_out; out=gpurun_out/ab_width.txt; : > $out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_width.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_width.log)" | tee -a $out
tail -30 gpurun_out/tests_width.log | grep -E "Error|assert|FAILED" | head -10 >> $out
for nb in "96 32" "48 32" "80 16" "112 8" "192 32" "256 256"; do
    set -- $nb
    for aw in 0 1; do
        HELMNET_TCF_ANY_WIDTH=$aw timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
            python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2 any_width=$aw', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'])
" >> $out 2>&1
    done
done
cat $out
