#!/bin/bash
# Round 2: packed narrow levels with the images per MMA picked by the step model (HELMNET_PACK_PENALTY), against off.
mkdir -p gpurun_out; out=gpurun_out/r2_pack2.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for rep in 1 2; do
HELMNET_PACK_NARROW=0 $q 256x256 256x32 96x32 64x32 128x64 256x64 --tag off >> $out 2>&1
HELMNET_PACK_PENALTY=12 $q 256x256 256x32 96x32 64x32 128x64 256x64 --tag pen12 >> $out 2>&1
HELMNET_PACK_PENALTY=4 $q 256x256 256x32 96x32 64x32 128x64 256x64 --tag pen4 >> $out 2>&1
HELMNET_PACK_PENALTY=25 $q 256x256 256x32 96x32 64x32 128x64 256x64 --tag pen25 >> $out 2>&1
done
cat $out
