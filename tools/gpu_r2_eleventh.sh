#!/bin/bash
# Round 2: full GPU suite + bench line (incl. the training_step extra) + training launch list after the weight-gradient rewrite.
mkdir -p gpurun_out; out=gpurun_out/r2_eleventh.txt; : > $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_eleventh.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_eleventh.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_eleventh.log | cut -c1-250 | head -20 >> $out
timeout 900 python bench.py > gpurun_out/bench_r2_eleventh.json 2> gpurun_out/bench_r2_eleventh.err
echo "bench rc=$?" | tee -a $out
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_eleventh.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/it", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "sustained", d["sustained"] and round(d["sustained"]["value"], 1))
print("training_step", d.get("training_step"))
print("roofline", d["roofline"]["frac"], "clocks", d["clocks"])
PY
timeout 300 python tools/train_probe.py 96 32 10 3 >> $out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_probe.py 96 32 1 2 > gpurun_out/ncu_train.log 2>&1
echo "ncu rc=$?" >> $out
cat $out
