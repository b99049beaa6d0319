#!/bin/bash
# One gpurun call: parity suite under the chosen programmatic-dependent-launch mode, then A/B timings of the HELMNET_PDL
# modes (0 off, 1 all kernels trigger early, 2 only one-CTA-per-SM tcgen05 kernels trigger, 3 no tcgen05 kernel triggers)
# on the headline config and the other BASELINE.json sizes.  Results -> gpurun_out/.
mkdir -p gpurun_out
out=gpurun_out/ab_pdl.txt
: > $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $out 2>&1
MODE=${1:-2}
HELMNET_PDL=$MODE timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_pdl$MODE.log 2>&1
echo "tests HELMNET_PDL=$MODE rc=$?  $(tail -1 gpurun_out/tests_pdl$MODE.log)" | tee -a $out
for nb in "256 256" "256 256" "512 64" "1024 8" "96 32" "256 1" "128 64"; do
    set -- $nb
    for p in 0 1 2 3; do
        HELMNET_PDL=$p timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
            python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2 pdl=$p', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'], 'clk', d['clocks']['sm_mhz'])
" >> $out 2>&1
    done
done
cat $out
