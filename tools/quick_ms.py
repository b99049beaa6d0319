"""ms per iteration of hn_run for a list of (n, batch) pairs under the current environment (A/B helper for gpurun scripts).

    python tools/quick_ms.py 256x256 256x32 96x32 [--iters 50] [--tag name]
Prints one line per pair: ms/iteration (CUDA events around hn_run), Mpoint-it/s, kernels per iteration, UNet / spectral stage ms.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CKPT, SOURCE_OF_N, make_maps  # noqa: E402
from helmnet_b200 import IterativeSolver  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if "x" in a and a.replace("x", "").isdigit()]
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 50
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else ""
    dev = torch.device("cuda", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for pair in args:
        n, b = (int(v) for v in pair.split("x"))
        s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
        s.freeze(); s.to(dev)
        s.set_domain_size(n, source_location=SOURCE_OF_N.get(n, [n // 8, n // 2]))
        base = make_maps(n, min(b, 32))
        sos = base.repeat((b + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:b].contiguous().to(dev)
        lib, ptr, st = s.lib, s._ptr, s._stream
        ctx = s._ensure_ctx(b)
        buf = torch.empty(iters, b, device=dev)
        lib.check(lib.hn_reset(ctx, ptr(sos), b, st()), "reset")
        lib.check(lib.hn_run(ctx, 5, ptr(buf), ptr(None), ptr(None), ptr(None), st()), "run")
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0.record()
            lib.check(lib.hn_run(ctx, iters, ptr(buf), ptr(None), ptr(None), ptr(None), st()), "run")
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters)
        stage = (C.c_float * 2)()
        acc = [0.0, 0.0]
        for _ in range(5):
            lib.check(lib.hn_profile_iteration(ctx, stage, st()), "prof")
            acc[0] += stage[0] / 5; acc[1] += stage[1] / 5
        s.sync_check()
        print(f"{tag:24s} n={n} batch={b}: {best:.4f} ms/it  {b * n * n / best / 1e3:8.1f} Mpoint-it/s  kernels {lib.hn_kernels_per_iteration(ctx)}"
              f"  unet {acc[0]:.4f} spectral {acc[1]:.4f}  rmse_last {float(buf[iters - 1].max()):.3e}", flush=True)
        s._release_ctx()
        del s
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
