#!/bin/bash
# ncu launch lists (per-kernel time and grid) of the other configurations' shares: where are grids underfilled?
mkdir -p gpurun_out
for cfg in 96x32 512x8 1024x1 256x64 256x128; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_list_$cfg.csv python tools/quick_ms.py $cfg --iters 3 > gpurun_out/ncu_list_$cfg.log 2>&1
    echo "$cfg rc=$?"
done
