#!/bin/bash
# Round 2: iterations per graph (HELMNET_GRAPH_UNROLL 1 / 2 / 4) at the per-GPU shares and small solves; full GPU suite.
mkdir -p gpurun_out; out=gpurun_out/r2_tenth.txt; : > $out
q="timeout 300 python tools/quick_ms.py"
for u in 1 2 4; do
HELMNET_GRAPH_UNROLL=$u $q 256x32 256x1 96x32 256x64 256x256 --tag unroll$u >> $out 2>&1
done
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_tenth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_tenth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_tenth.log | cut -c1-250 | head -20 >> $out
cat $out
