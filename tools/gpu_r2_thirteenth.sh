#!/bin/bash
# Round 2: balanced strips also in the down- / up-sampling kernels (HELMNET_DCONV_BALANCE 1 / 2); full GPU suite.
mkdir -p gpurun_out; out=gpurun_out/r2_thirteenth.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for u in 1 2; do
HELMNET_DCONV_BALANCE=$u $q 256x32 256x64 256x128 256x256 96x32 128x64 512x8 1024x1 --tag balance$u >> $out 2>&1
done
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_thirteenth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_thirteenth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_thirteenth.log | cut -c1-250 | head -20 >> $out
cat $out
