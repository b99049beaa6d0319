#!/bin/bash
# One gpurun call: parity suite with programmatic dependent launch (HELMNET_PDL=1), a short subset with it off, then
# A/B timings (PDL x spectral chunking) on the headline config and the other BASELINE.json sizes.  Results -> gpurun_out/.
mkdir -p gpurun_out
out=gpurun_out/ab_pdl.txt
: > $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $out 2>&1

HELMNET_PDL=1 timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/tests_pdl1.log 2>&1
echo "tests HELMNET_PDL=1 rc=$?  $(tail -1 gpurun_out/tests_pdl1.log)" | tee -a $out
HELMNET_PDL=0 timeout 300 python -m pytest tests -m gpu -x -q -k "readme or n96_golden or single_step or launch_accounting" > gpurun_out/tests_pdl0.log 2>&1
echo "tests HELMNET_PDL=0 (subset) rc=$?  $(tail -1 gpurun_out/tests_pdl0.log)" | tee -a $out

for cfg in "0 0" "1 0" "0 0" "1 0" "1 64" "1 32"; do
    set -- $cfg
    HELMNET_PDL=$1 HELMNET_SPEC_CHUNK=$2 timeout 300 tools/bench_kernels.sh "pdl=$1 chunk=$2" >> $out 2>&1
done
for nb in "96 32" "512 64" "1024 8" "256 1"; do
    set -- $nb
    for p in 0 1; do
        HELMNET_PDL=$p timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
            python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2 pdl=$p', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'])
" >> $out 2>&1
    done
done
HELMNET_PDL=1 timeout 600 python bench.py > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err
echo "full bench pdl=1 rc=$?" >> $out
cat $out
