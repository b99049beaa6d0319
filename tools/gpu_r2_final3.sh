#!/bin/bash
# Round 2, last build: bench line (default flags, as the driver runs it) + the same with --steps 20 --warmup 5 (the driver's r1 choice).
mkdir -p gpurun_out; out=gpurun_out/r2_final3.txt; : > $out
timeout 900 python bench.py > gpurun_out/bench_r2_final5.json 2> gpurun_out/bench_r2_final5.err
echo "bench rc=$?" | tee -a $out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-extras --residual-iters 0 > gpurun_out/bench_r2_final5_k20.json 2>> gpurun_out/bench_r2_final5.err
echo "bench K=20 rc=$?" | tee -a $out
python - >> $out 2>&1 <<'PY'
import json
for f in ("gpurun_out/bench_r2_final5.json", "gpurun_out/bench_r2_final5_k20.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "steps", d["steps"], "value", round(d["value"], 1), "ms/it", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "sustained", d["sustained"] and round(d["sustained"]["value"], 1), "kernels", d["kernels_per_iteration"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    if d.get("other_configs"):
        print("  others", json.dumps({k: (round(v["ms_per_step"], 4), round(v["value"], 1)) for k, v in d["other_configs"].items()}))
        print("  readme", d.get("readme_lens_ms_to_residual_1e-3")); print("  training", d.get("training_step"))
        print("  cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"]["kind"]), "gpu_eager", d["gpu_eager_baseline"] and d["gpu_eager_baseline"].get("value"))
    for r in [d["roofline"]] + d["roofline_kernels"]:
        print(f"    {r.get('kernel', '')[:58]:58s} {r['achieved']:7.0f} GB/s  {r['frac']:.3f}")
    print("  stage unet", round(d["roofline_stage_unet"]["stage_ms"], 4), round(d["roofline_stage_unet"]["frac"], 4), "spectral", round(d["roofline_stage_spectral"]["stage_ms"], 4), round(d["roofline_stage_spectral"]["frac"], 4))
PY
cat $out
