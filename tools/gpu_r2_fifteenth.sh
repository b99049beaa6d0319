#!/bin/bash
# Round 2: balanced strips in the per-conv kernel (level 0 of the 512^2 / 1024^2 domains): HELMNET_DCONV_BALANCE 1 / 2; full GPU suite.
mkdir -p gpurun_out; out=gpurun_out/r2_fifteenth.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for u in 1 2; do
HELMNET_DCONV_BALANCE=$u $q 512x8 1024x1 512x64 1024x8 512x3 --iters 30 --tag balance$u >> $out 2>&1
done
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_fifteenth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_fifteenth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_fifteenth.log | cut -c1-250 | head -20 >> $out
cat $out
