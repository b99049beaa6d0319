#!/bin/bash
# One gpurun call: GPU parity suite, the default bench line, a few other sizes, smoke().  Results -> gpurun_out/final2.txt
mkdir -p gpurun_out; out=gpurun_out/final2.txt; : > $out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/tests_final.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_final.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_final.log | head -20 >> $out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?" | tee -a $out
for nb in "96 32" "128 64" "256 1" "512 64" "1024 8"; do
    set -- $nb
    timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
        python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'], 'roofline', d['roofline']['kernel'][:40], round(d['roofline']['frac'], 3))
" >> $out 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" >> $out 2>&1
timeout 200 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1
cat $out
