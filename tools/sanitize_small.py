"""Small solves for compute-sanitizer (memcheck): every engine, fused and unfused levels, side branch, per-sample sources.

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CKPT  # noqa: E402
from helmnet_b200 import IterativeSolver  # noqa: E402

s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
s.freeze()
s.to("cuda:0")
g = torch.Generator().manual_seed(0)
for n, b, iters in ((64, 3, 3), (96, 2, 2), (256, 2, 2), (512, 1, 1)):
    for eng in (2, 1, 0):
        if eng == 0 and n > 96:
            continue
        s.set_engine(eng)
        s.set_domain_size(n, source_location=[n // 8, n // 2])
        if n == 64:
            s.set_multiple_sources([[5, 6], [20, 30], [40, 50]])
        sos = (1.0 + 0.5 * torch.rand(b, 1, n, n, generator=g)).cuda()
        out = s.forward(sos, num_iterations=iters, return_wavefields=True, return_states=True)
        s.sync_check()
        assert torch.isfinite(out["wavefields"][-1]).all()
        print(f"n={n} b={b} engine={eng}: ok, rmse {out['residual_rmse'][-1].tolist()}", flush=True)
# batch sizes that take the balanced strips (chunks across image boundaries) and the packed narrow levels (several images per MMA, partly
# filled last group), engine 2 only; HELMNET_PACK_NARROW=2 in the environment also packs the fused DoubleConv kernels
s.set_engine(2)
for n, b, iters in ((64, 41, 2), (96, 37, 2), (256, 37, 1), (128, 75, 1)):
    s.set_domain_size(n, source_location=[n // 8, n // 2])
    sos = (1.0 + 0.5 * torch.rand(b, 1, n, n, generator=g)).cuda()
    out = s.forward(sos, num_iterations=iters)
    s.sync_check()
    assert torch.isfinite(out["wavefields"][-1]).all()
    print(f"n={n} b={b} engine=2 (balanced / packed): ok", flush=True)
print("done")
