#!/bin/bash
# Round 2: threshold of the balanced-strip choice (HELMNET_BAL_THRESH 97 / 100 / 104 per cent of the uniform cost), interleaved.
mkdir -p gpurun_out; out=gpurun_out/r2_thresh.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for rep in 1 2; do
for u in 97 100 104; do
HELMNET_BAL_THRESH=$u $q 256x32 256x64 256x16 96x32 128x64 512x8 --tag thr$u >> $out 2>&1
done
done
cat $out
