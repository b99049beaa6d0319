#!/bin/bash
# Round 2: A/B of the per-launch strip height of the down-sampling kernel (HELMNET_DOWN_ROWS=32 = the fixed height of r1), parity suite.
mkdir -p gpurun_out; out=gpurun_out/r2_sixth.txt; : > $out
q="timeout 300 python tools/quick_ms.py"
$q 256x256 256x128 256x64 256x32 256x16 256x8 256x1 96x32 128x64 64x256 512x8 --tag auto_rows >> $out 2>&1
HELMNET_DOWN_ROWS=32 $q 256x256 256x128 256x64 256x32 256x16 256x8 256x1 96x32 128x64 64x256 512x8 --tag rows32 >> $out 2>&1
HELMNET_DOWN_ROWS=64 $q 256x256 256x128 --tag rows64 >> $out 2>&1
$q 256x256 256x32 --tag auto_again >> $out 2>&1
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_sixth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_sixth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_sixth.log | cut -c1-300 | head -30 >> $out
cat $out
