#!/bin/bash
# Round 2, final single-GPU evidence of the build with balanced strips: bench line, reference arm, ncu launch list + --set full capture.
mkdir -p gpurun_out; out=gpurun_out/r2_final2.txt; : > $out
timeout 900 python bench.py > gpurun_out/bench_r2_final4.json 2> gpurun_out/bench_r2_final4.err
echo "bench rc=$?" | tee -a $out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final4_reference.json 2>> gpurun_out/bench_r2_final4.err
echo "bench reference rc=$?" | tee -a $out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $out
timeout 420 ncu --set full --clock-control none --import-source on -k regex:"dconv_tcf|spectral|down_tcr|up_tcr" -s 14 -c 14 \
    -o gpurun_out/prof_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $out
ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>> $out
rm -f gpurun_out/prof_r2.ncu-rep
python - >> $out 2>&1 <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final4.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/it", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "sustained", d["sustained"] and round(d["sustained"]["value"], 1), "kernels", d["kernels_per_iteration"])
print("cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"]["kind"]), "gpu_eager", d["gpu_eager_baseline"] and d["gpu_eager_baseline"].get("value"))
print("others", json.dumps({k: (round(v["ms_per_step"], 4), round(v["value"], 1)) for k, v in (d.get("other_configs") or {}).items()}))
print("readme", d.get("readme_lens_ms_to_residual_1e-3")); print("training", d.get("training_step"))
for r in [d["roofline"]] + d["roofline_kernels"]:
    print(f"  {r.get('kernel', '')[:60]:60s} {r.get('us', 0):8.1f} us {r['achieved']:7.0f} GB/s  {r['frac']:.3f}")
print("stage unet", d["roofline_stage_unet"]["stage_ms"], d["roofline_stage_unet"]["frac"], "spectral", d["roofline_stage_spectral"]["stage_ms"], d["roofline_stage_spectral"]["frac"])
print("clocks", d["clocks"])
PY
cat $out
