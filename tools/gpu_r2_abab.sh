#!/bin/bash
# Interleaved A/B/A/B of the balanced strips at the headline size (order effects: clocks / temperature drift within a call).
mkdir -p gpurun_out; out=gpurun_out/r2_abab.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for u in 0 2 0 2 0 2; do
HELMNET_DCONV_BALANCE=$u $q 256x256 256x32 --iters 100 --tag balance$u >> $out 2>&1
done
cat $out
