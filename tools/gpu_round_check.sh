#!/bin/bash
# One gpurun call: GPU parity suite, the default bench line, smoke().  Results -> gpurun_out/final2.txt
mkdir -p gpurun_out; out=gpurun_out/final2.txt; : > $out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/tests_final.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_final.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_final.log | head -20 >> $out
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?" | tee -a $out
python -c "import __graft_entry__ as g; g.smoke()" >> $out 2>&1
cat $out
