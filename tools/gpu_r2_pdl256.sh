#!/bin/bash
# PDL / side branch at the headline size now that the strips are balanced (all CTAs of a kernel finish together): interleaved A/B.
mkdir -p gpurun_out; out=gpurun_out/r2_pdl256.txt; : > $out
q="timeout 400 python tools/quick_ms.py"
for rep in 1 2; do
$q 256x256 --iters 50 --tag default >> $out 2>&1
HELMNET_PDL=2 $q 256x256 --iters 50 --tag pdl2 >> $out 2>&1
HELMNET_PDL=1 $q 256x256 --iters 50 --tag pdl1 >> $out 2>&1
HELMNET_PDL=2 HELMNET_SIDE_STATE=1 $q 256x256 --iters 50 --tag pdl2_side >> $out 2>&1
HELMNET_PDL=0 HELMNET_SIDE_STATE=1 $q 256x256 --iters 50 --tag side >> $out 2>&1
done
cat $out
