#!/bin/bash
# Round 2: balanced strips of the fused DoubleConv kernels (HELMNET_DCONV_BALANCE 0 / 1) at the per-GPU shares; full GPU suite.
mkdir -p gpurun_out; out=gpurun_out/r2_twelfth.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for u in 0 1; do
HELMNET_DCONV_BALANCE=$u $q 256x32 256x64 256x128 256x256 256x16 96x32 128x64 512x8 --tag balance$u >> $out 2>&1
done
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_twelfth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_twelfth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_twelfth.log | cut -c1-250 | head -20 >> $out
cat $out
