#!/bin/bash
# One short gpurun call: GPU parity suite + A/B of this build against a baseline library (tools/libhelmnet_base.so, built by hand
# with the option under test switched off) on the headline config and a few smaller sizes.
mkdir -p gpurun_out; out=gpurun_out/quick.txt; : > $out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/tests_quick.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_quick.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_quick.log | head -20 >> $out
for nb in "256 256" "64 256" "32 256" "96 32" "256 256"; do
    set -- $nb
    for lib in base new; do
        if [ $lib = base ]; then export HELMNET_SM100_LIB=$PWD/tools/libhelmnet_base.so; else unset HELMNET_SM100_LIB; fi
        timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
            python -c "
import sys, json
d = json.loads(sys.stdin.read())
ks = [d['roofline']] + d['roofline_kernels']
print('n=$1 batch=$2 $lib', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'], '| us:', ' '.join('%.1f' % (r['ms_per_launch'] * 1e3) for r in ks[:4]))
" >> $out 2>&1
    done
done
cat $out
