#!/bin/bash
# Round 2: per-launch strip height of the per-conv kernel (1024^2 / 512^2 shares), knobs for the 96^2 configuration.
mkdir -p gpurun_out; out=gpurun_out/r2_ninth.txt; : > $out
q="timeout 300 python tools/quick_ms.py"
$q 1024x1 512x8 512x1 96x32 256x256 --tag base >> $out 2>&1
HELMNET_TCD_MIN_RES=2 $q 96x32 48x32 --tag tcd2 >> $out 2>&1
HELMNET_TCF_MIN_WIDTH=6 $q 96x32 --tag tcfmin6 >> $out 2>&1
HELMNET_SPEC_L=8 HELMNET_SPEC_CW=8 $q 96x32 --tag spec8 >> $out 2>&1
HELMNET_SPEC_L=4 HELMNET_SPEC_CW=4 $q 96x32 --tag spec4 >> $out 2>&1
HELMNET_TCD_MIN_RES=2 HELMNET_TCF_MIN_WIDTH=6 HELMNET_SPEC_L=8 HELMNET_SPEC_CW=8 $q 96x32 --tag all >> $out 2>&1
HELMNET_TCD_MIN_RES=2 HELMNET_TCF_MIN_WIDTH=6 timeout 900 python -m pytest tests -m gpu -q -k "forward_vs_oracle or n96 or layers or srcmap or multiple_sources or test_step" > gpurun_out/tests_r2_ninth_env.log 2>&1
echo "tests (TCD_MIN_RES=2, TCF_MIN_WIDTH=6) rc=$?  $(tail -1 gpurun_out/tests_r2_ninth_env.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_ninth_env.log | cut -c1-250 | head -20 >> $out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_ninth.log 2>&1
echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_ninth.log)" | tee -a $out
grep -E "^FAILED|^E  " gpurun_out/tests_r2_ninth.log | cut -c1-250 | head -20 >> $out
cat $out
