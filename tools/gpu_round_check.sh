#!/bin/bash
# What the driver runs at round end: GPU suite, smoke(), default bench line.
mkdir -p gpurun_out; out=gpurun_out/round_check.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_round_check.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_round_check.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_round_check.log | cut -c1-250 | head -20 >> $out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_round_check.log 2>&1
echo "smoke rc=$?" | tee -a $out
grep smoke gpurun_out/smoke_round_check.log >> $out
tail -3 gpurun_out/smoke_round_check.log | cut -c1-300 >> $out
cat $out
