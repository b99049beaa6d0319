"""Generate tests/golden/* by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY -- run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference (helmnet/, pure Python) is imported from /root/reference with the
three stub packages in oracle/_stubs standing in for pytorch_lightning,
torchmetrics and matplotlib (not installed in this image).
No reference source is copied: only tensors it computes are stored.

Fixtures written (float32 unless noted):
  jcp_paper_trained_weights_slim.ckpt  f.* weights + source + hparams of the shipped ckpt, re-saved
                                       as a plain-dict zip checkpoint (no Lightning classes inside)
  operator_n{32,96}.npz                PML/k tables and L(u) for a seeded random u
  unet_step_n32.npz                    one HybridNet.forward + single_step from random inputs
  traj_n96_b2.npz                      forward(), 40 iterations, 2 synthetic sos maps
  traj_readme_n256.npz                 README lens example, 120 iterations
  traj_srcmap_n64.npz                  examples/simple_scattering.py style source map, 64^2
  traj_bench_n256_b2.npz               bench.py's workload (first two synthetic 256^2 maps), 12 iterations
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HELMNET_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, REF)
sys.path.insert(2, os.path.dirname(HERE))

from helmnet import IterativeSolver  # noqa: E402  (the reference)
from helmnet_b200.synthetic import synthetic_sos  # noqa: E402  (input generator only)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CKPT = os.path.join(REF, "trained_models", "jcp_paper_trained_weights.ckpt")


def load_solver():
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    s.freeze()
    return s


def np32(t):
    return t.detach().contiguous().cpu().numpy().astype(np.float32)


def bench_workload_fixture(solver):
    """traj_bench_n256_b2.npz: the first two maps of bench.py's workload (synthetic_sos(.., 256, seed=1), point source [30,128]),
    12 iterations of the unmodified reference."""
    n, b, iters = 256, 2, 12
    with torch.no_grad():
        solver.hparams.source_location = [30, 128]
        solver.set_domain_size(n, source_location=[30, 128])
        sos = synthetic_sos(32, n, seed=1)[:b].contiguous()
        out = solver.forward(sos, num_iterations=iters, return_wavefields=True)
        rm = torch.stack([solver.test_loss_function(r) for r in out["residuals"]], 0)
        keep = [0, iters - 1]
        np.savez_compressed(
            os.path.join(OUT, "traj_bench_n256_b2.npz"),
            sos=np32(sos), rmse=np32(rm), keep=np.array(keep),
            wavefields=np.stack([np32(out["wavefields"][i]) for i in keep]),
            source_location=np.array([30, 128]),
        )


def main():
    if "--only-bench" in sys.argv:      # add the bench-workload fixture without rewriting the others
        torch.manual_seed(0)
        np.random.seed(0)
        torch.set_num_threads(os.cpu_count())
        bench_workload_fixture(load_solver())
        print("written", os.path.join(OUT, "traj_bench_n256_b2.npz"))
        return
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count())
    solver = load_solver()

    # ---- slim checkpoint -------------------------------------------------
    raw = torch.load(CKPT, map_location="cpu", weights_only=False)
    sd = {k: v.clone() for k, v in raw["state_dict"].items() if k.startswith("f.") or k == "source"}
    hp = {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in dict(raw["hyper_parameters"]).items()}
    torch.save({"hyper_parameters": hp, "state_dict": sd, "pytorch-lightning_version": raw["pytorch-lightning_version"],
                "slim_of": "trained_models/jcp_paper_trained_weights.ckpt"},
               os.path.join(OUT, "jcp_paper_trained_weights_slim.ckpt"))
    with open(os.path.join(OUT, "hparams.json"), "w") as f:
        json.dump(hp, f, indent=1, sort_keys=True)

    with torch.no_grad():
        # ---- operator KATs ----------------------------------------------
        for n in (32, 96):
            # the reference rebuilds its default source at hparams.source_location inside
            # set_domain_size (hybridnet.py:96,133-143) and raises IndexError if that lies outside
            # the new domain, so move it first for small domains
            solver.hparams.source_location = [n // 3, n // 2]
            solver.set_domain_size(n, source_location=[n // 3, n // 2])
            lap = solver.Lap
            g = torch.Generator().manual_seed(100 + n)
            u = torch.randn(2, n, n, 2, generator=g)
            np.savez_compressed(
                os.path.join(OUT, f"operator_n{n}.npz"),
                u=np32(u), Lu=np32(lap(u)),
                ax=np32(lap.ax[0, 0, :, :]), bx=np32(lap.bx[0, 0, :, :]),     # vary along W
                ay=np32(lap.ay[0, :, 0, :]), by=np32(lap.by[0, :, 0, :]),     # vary along H
                kx=np32(lap.kx[0, 0, :, 1]), kx_sq=np32(lap.kx_sq[0, 0, :, 0]),
                sigmas=np32(solver.sigmas), source=np32(solver.source),
            )

        # ---- one UNet call + one solver step from random state ---------------
        n, b = 32, 2
        solver.hparams.source_location = [10, 16]
        solver.set_domain_size(n, source_location=[10, 16])
        g = torch.Generator().manual_seed(7)
        wf = torch.randn(b, 2, n, n, generator=g) * 0.5
        res = torch.randn(b, 2, n, n, generator=g) * 5e-3
        sos = 1.0 + torch.rand(b, 1, n, n, generator=g)
        k_sq = (solver.hparams.omega / sos) ** 2
        states = [torch.randn(b, 2, n // 2 ** d, n // 2 ** d, generator=g) * 0.3 for d in range(4)]
        solver.f.set_states(torch.cat([s.reshape(b, 2, -1) for s in states], 2), flatten=True)
        sig = solver.sigmas.unsqueeze(0).repeat(b, 1, 1, 1)
        inp = torch.cat([wf, 1e3 * res, sig], 1)
        d_wf = solver.f(inp)
        st1 = solver.f.get_states(flatten=True)
        # second call continues from the new state through single_step
        solver.f.set_states(torch.cat([s.reshape(b, 2, -1) for s in states], 2), flatten=True)
        up, new_res = solver.single_step(wf, k_sq, res)
        np.savez_compressed(
            os.path.join(OUT, "unet_step_n32.npz"),
            wf=np32(wf), res=np32(res), sos=np32(sos), k_sq=np32(k_sq),
            states_flat=np32(torch.cat([s.reshape(b, 2, -1) for s in states], 2)),
            inp=np32(inp), d_wf=np32(d_wf), states_flat_out=np32(st1),
            up_wf=np32(up), new_res=np32(new_res), source=np32(solver.source),
            residual_of_wf=np32(solver.get_residual(wf, k_sq)),
        )

        # ---- trajectory, 96^2, batch 2 -------------------------------------
        n, b, iters = 96, 2, 40
        solver.hparams.source_location = [82, 48]
        solver.set_domain_size(n, source_location=[82, 48])
        sos = synthetic_sos(b, n, seed=0)
        out = solver.forward(sos, num_iterations=iters, return_wavefields=True, return_states=True)
        rm = torch.stack([solver.test_loss_function(r) for r in out["residuals"]], 0)
        keep = [0, 1, 9, iters - 1]
        np.savez_compressed(
            os.path.join(OUT, "traj_n96_b2.npz"),
            sos=np32(sos), rmse=np32(rm), keep=np.array(keep),
            wavefields=np.stack([np32(out["wavefields"][i]) for i in keep]),
            residuals=np.stack([np32(out["residuals"][i]) for i in keep]),
            states_last=np32(out["states"][-1]), source=np32(solver.source),
            source_location=np.array([82, 48]),
        )

        # ---- README lens, 256^2 --------------------------------------------
        n, iters = 256, 120
        sos_map = np.ones((n, n))
        sos_map[100:170, 30:240] = np.tile(np.linspace(2, 1, 210), (70, 1))
        solver.hparams.source_location = [30, 128]
        solver.set_domain_size(n, source_location=[30, 128])
        sos = torch.tensor(sos_map).float()[None, None]
        out = solver.forward(sos, num_iterations=iters, return_wavefields=True)
        rm = torch.stack([solver.test_loss_function(r) for r in out["residuals"]], 0)
        keep = [0, 9, 49, iters - 1]
        np.savez_compressed(
            os.path.join(OUT, "traj_readme_n256.npz"),
            rmse=np32(rm), keep=np.array(keep),
            wavefields=np.stack([np32(out["wavefields"][i]) for i in keep]),
            source_location=np.array([30, 128]),
        )

        # ---- source map (examples/simple_scattering.py), scaled to 64^2, batch 3 with per-sample maps
        n, iters = 64, 30
        src = np.zeros((3, 2, n, n), np.float32)
        src[0, 0, 8, 28:36] = 1
        src[1, 0, 50, 10:14] = 1
        src[1, 1, 50, 10:14] = 0.5
        src[2, 1, 20:24, 40] = -1
        sos_map = np.ones((3, 1, n, n), np.float32)
        sos_map[:, 0, 25:42, 8:60] = 1.5
        sos_map[2, 0, 30:35, 30:50] = 1.9
        solver.hparams.source_location = [10, 10]
        solver.set_domain_size(n, source_map=torch.tensor(src))
        out = solver.forward(torch.tensor(sos_map), num_iterations=iters, return_wavefields=False)
        rm = torch.stack([solver.test_loss_function(r) for r in out["residuals"]], 0)
        np.savez_compressed(
            os.path.join(OUT, "traj_srcmap_n64.npz"),
            sos=sos_map, source=src, rmse=np32(rm), wavefield=np32(out["wavefields"][0]),
            residual=np32(out["residuals"][-1]),
        )
    bench_workload_fixture(solver)
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn:45s} {os.path.getsize(os.path.join(OUT, fn)) / 1024:8.1f} KB")


if __name__ == "__main__":
    main()
