#!/bin/bash
# Round 2, SURVEY 8(f4): gradient parity of the training unroll on the B200 + time of one training step + launch list of one step.
mkdir -p gpurun_out; out=gpurun_out/r2_train.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -k "training" > gpurun_out/tests_r2_train.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_train.log)" | tee -a $out
grep -E "^FAILED|^E  |training step" gpurun_out/tests_r2_train.log | cut -c1-300 | head -30 >> $out
timeout 300 python tools/train_probe.py 96 32 10 3 >> $out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_probe.py 96 32 1 2 > gpurun_out/ncu_train.log 2>&1
echo "ncu rc=$?" >> $out
cat $out
