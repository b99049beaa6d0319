#!/bin/bash
# One gpurun call: parity suite with the defaults and with shorter fused-DoubleConv strips (HELMNET_DCONV_MIN_ROWS), then
# small-batch latency A/B.  Results -> gpurun_out/ab_rows.txt
mkdir -p gpurun_out
out=gpurun_out/ab_rows.txt
: > $out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_default.log 2>&1
echo "tests defaults rc=$?  $(tail -1 gpurun_out/tests_default.log)" | tee -a $out
for mr in 2 4; do
    HELMNET_DCONV_MIN_ROWS=$mr timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_minrows$mr.log 2>&1
    echo "tests HELMNET_DCONV_MIN_ROWS=$mr rc=$?  $(tail -1 gpurun_out/tests_minrows$mr.log)" | tee -a $out
done
for nb in "256 1" "256 2" "256 8" "256 32" "128 4" "256 256"; do
    set -- $nb
    for mr in 8 4 2; do
        HELMNET_DCONV_MIN_ROWS=$mr timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
            python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2 min_rows=$mr', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'])
" >> $out 2>&1
    done
done
cat $out
