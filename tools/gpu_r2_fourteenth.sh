#!/bin/bash
# Round 2: bit-identity of the balanced strips; ncu launch lists of the current build (256^2 x 256 and the 32-map share).
mkdir -p gpurun_out; out=gpurun_out/r2_fourteenth.txt; : > $out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "balanced or scheduling or strip_paths" > gpurun_out/tests_r2_fourteenth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_fourteenth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_fourteenth.log | cut -c1-250 | head -20 >> $out
for cfg in 256x256 256x32; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_list_$cfg.csv python tools/quick_ms.py $cfg --iters 3 > gpurun_out/ncu_list_$cfg.log 2>&1
    echo "$cfg ncu rc=$?" >> $out
done
cat $out
