#!/usr/bin/env python
"""gpurun_out/parity_measured.jsonl (written by tests/test_gpu_parity.py on the GPU box) -> profiles/<round>_parity_report.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl")
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"
out = [f"# Measured parity margins of the GPU suite (tests/test_gpu_parity.py -> gpurun_out/parity_measured.jsonl), one B200, round {rnd[1:]}.",
       "# engine 0 = fp32 CUDA cores, 1 = tcgen05 one kernel per conv, 2 = tcgen05 fused DoubleConvs (default).  rel-L2 unless stated.",
       "# '*_vs_fp64' = against the reference's float64 run (the arbiter), 'ref32_vs_fp64' = the reference's own fp32 run against it.",
       "# Bars (BASELINE.json): per-iteration wavefield 1e-5, final wavefield 1e-3, residual-norm trajectory within the same bounds."]
for line in open(src):
    d = json.loads(line)
    t = d.pop("test")
    out.append(f"{t:30s} " + "  ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in d.items()))
open(os.path.join(ROOT, "profiles", f"{rnd}_parity_report.txt"), "w").write("\n".join(out) + "\n")
print(len(out) - 4, "records")
