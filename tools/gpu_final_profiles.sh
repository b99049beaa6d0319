#!/bin/bash
# One gpurun call at the end of a round: GPU parity suite, the full bench line, the ncu launch list and --set full capture
# of the same command (tools/make_profiles.py turns them into profiles/*), parity margins, and the other BASELINE.json sizes.
mkdir -p gpurun_out
out=gpurun_out/final.txt
: > $out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_final.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_final.log)" | tee -a $out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?" | tee -a $out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
echo "bench reference rc=$?" | tee -a $out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --residual-iters 0 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $out
timeout 420 ncu --set full --clock-control none --import-source on -k regex:"dconv_tcf|spectral|down_tcr|up_tcr" -s 14 -c 14 \
    -o gpurun_out/prof_e2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --residual-iters 0 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $out
ncu -i gpurun_out/prof_e2.ncu-rep --page raw --csv > gpurun_out/prof_e2_raw.csv 2>> $out
rm -f gpurun_out/prof_e2.ncu-rep
timeout 200 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1
echo "parity report rc=$?" | tee -a $out
for nb in "96 32" "512 64" "1024 8" "256 1"; do
    set -- $nb
    timeout 300 python bench.py --n $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null |
        python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=$1 batch=$2', 'ms/it', round(d['ms_per_step'], 4), 'Mpoint-it/s', round(d['value'], 1), 'kernels', d['kernels_per_iteration'])
" >> $out 2>&1
done
timeout 120 python tools/latency_probe.py 2>&1 | grep "engine 2" >> $out
cat $out
