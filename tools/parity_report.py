#!/usr/bin/env python
"""Prints the measured parity margins of the CUDA path against the reference's golden trajectories (tests/golden) for each
conv engine: the numbers the -m gpu parity tests assert on.  Run on a GPU box: python tools/parity_report.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from helmnet_b200 import IterativeSolver  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def main():
    s = IterativeSolver.load_from_checkpoint(os.path.join(GOLD, "jcp_paper_trained_weights_slim.ckpt"), strict=False, test_data_path=None)
    s.freeze()
    s.to("cuda:0")
    print("# parity of the CUDA path vs the golden trajectories produced by the unmodified reference (CPU fp32)")
    print("# bars (BASELINE.json): per-iteration wavefield rel-L2 <= 1e-5, final <= 1e-3, residual-norm trajectory <= 1e-5 (1e-4 README)")
    for eng, name in ((2, "tcgen05 fused DoubleConv"), (1, "tcgen05 per conv"), (0, "fp32 CUDA cores")):
        s.set_engine(eng)
        # README lens, 256^2, 120 iterations
        g = np.load(os.path.join(GOLD, "traj_readme_n256.npz"))
        lens = np.ones((256, 256), np.float32)
        lens[100:170, 30:240] = np.tile(np.linspace(2, 1, 210), (70, 1))
        s.set_domain_size(256, source_location=[30, 128])
        with torch.no_grad():
            out = s.forward(torch.from_numpy(lens)[None, None].cuda(), num_iterations=120, return_wavefields=True)
        rm = out["residual_rmse"].cpu().numpy()[:, 0]
        e_rm = float(np.max(np.abs(rm - g["rmse"][:, 0]) / g["rmse"][:, 0]))
        errs = ", ".join(f"it {int(k)}: {rel(out['wavefields'][int(k)], g['wavefields'][i]):.2e}" for i, k in enumerate(g["keep"]))
        print(f"engine {eng} ({name}) README lens 256^2: wavefield rel-L2 {errs}; max rel. RMSE-trajectory error {e_rm:.2e}; "
              f"first RMSE<1e-3 at {int(np.argmax(rm < 1e-3))}")
        # 96^2 x 2, 40 iterations
        g = np.load(os.path.join(GOLD, "traj_n96_b2.npz"))
        s.set_domain_size(96, source_location=[int(v) for v in g["source_location"]] if "source_location" in g.files else [82, 48])
        with torch.no_grad():
            out = s.forward(torch.tensor(g["sos"]).cuda(), num_iterations=int(g["rmse"].shape[0]), return_wavefields=True)
        ew = ", ".join(f"it {int(k)}: {rel(out['wavefields'][int(k)], g['wavefields'][i]):.2e}" for i, k in enumerate(g["keep"]))
        print(f"engine {eng} ({name}) 96^2 x 2: wavefield rel-L2 {ew}; RMSE trajectory rel-L2 {rel(out['residual_rmse'], g['rmse']):.2e}")


if __name__ == "__main__":
    main()
