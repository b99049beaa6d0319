#!/bin/bash
# Round 2, last build: launch knobs at the 32-map share once more (PDL modes, strip floor of the fused kernels), interleaved.
mkdir -p gpurun_out; out=gpurun_out/r2_knobs2.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for rep in 1 2; do
$q 256x32 96x32 --tag default >> $out 2>&1
HELMNET_PDL=1 $q 256x32 96x32 --tag pdl1 >> $out 2>&1
HELMNET_PDL=3 $q 256x32 96x32 --tag pdl3 >> $out 2>&1
HELMNET_DCONV_MIN_ROWS=2 $q 256x32 96x32 --tag minrows2 >> $out 2>&1
HELMNET_DCONV_MIN_ROWS=8 $q 256x32 96x32 --tag minrows8 >> $out 2>&1
HELMNET_SPEC_L=4 HELMNET_SPEC_CW=4 $q 96x32 --tag spec4 >> $out 2>&1
done
cat $out
