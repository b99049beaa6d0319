#!/bin/bash
# Round 2: N = 256 column kernel with the operator tables read through L1 (35 KB shared memory, 6 CTAs per SM): HELMNET_COLS_LEAN 0 / 1.
mkdir -p gpurun_out; out=gpurun_out/r2_lean.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for u in 0 1 0 1; do
HELMNET_COLS_LEAN=$u $q 256x256 256x32 256x64 256x1 --tag lean$u >> $out 2>&1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "laplacian or readme_golden or residual" > gpurun_out/tests_r2_lean.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_lean.log)" | tee -a $out
cat $out
