#!/bin/bash
# Round 2: packed narrow levels also in the fused DoubleConv kernels (HELMNET_PACK_NARROW 1 / 2); full GPU suite with the default (2).
mkdir -p gpurun_out; out=gpurun_out/r2_pack3.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_pack3.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_pack3.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_pack3.log | cut -c1-250 | head -20 >> $out
q="timeout 400 python tools/quick_ms.py"
for u in 1 2 1 2; do
HELMNET_PACK_NARROW=$u $q 256x256 256x32 96x32 64x32 128x64 256x64 --tag pack$u >> $out 2>&1
done
cat $out
