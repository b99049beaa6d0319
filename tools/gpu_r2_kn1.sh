#!/bin/bash
# Round 2: conv_state's first conv with the horizontal taps of its 8-channel group in N (HELMNET_KN1 0 / 1): 5 -> 3 MMAs per 128 pixels.
mkdir -p gpurun_out; out=gpurun_out/r2_kn1.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_kn1.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_kn1.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_kn1.log | cut -c1-250 | head -20 >> $out
q="timeout 400 python tools/quick_ms.py"
for u in 0 1 0 1; do
HELMNET_KN1=$u $q 256x256 256x128 256x64 128x64 --tag kn$u >> $out 2>&1
done
cat $out
