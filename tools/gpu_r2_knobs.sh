#!/bin/bash
# Round 2: knob A/B at the small shares with the final build (no code change).
mkdir -p gpurun_out; out=gpurun_out/r2_knobs.txt; : > $out
q="timeout 300 python tools/quick_ms.py"
$q 256x32 256x64 256x128 96x32 --tag base >> $out 2>&1
HELMNET_DCONV_MIN_ROWS=2 $q 256x32 256x64 96x32 --tag minrows2 >> $out 2>&1
HELMNET_DCONV_MIN_ROWS=8 $q 256x32 256x64 96x32 --tag minrows8 >> $out 2>&1
HELMNET_PDL=1 $q 256x32 256x64 256x128 96x32 --tag pdl1 >> $out 2>&1
HELMNET_PDL=2 HELMNET_SIDE_STATE=1 $q 256x128 256x256 --tag pdl2_side1_big >> $out 2>&1
HELMNET_PDL=0 $q 256x32 --tag pdl0 >> $out 2>&1
HELMNET_SIDE_STATE=0 $q 256x32 256x64 --tag side0 >> $out 2>&1
HELMNET_TCF_MIN_WIDTH=32 $q 256x32 --tag tcfmin32 >> $out 2>&1
$q 256x32 256x64 --tag base_again >> $out 2>&1
cat $out
