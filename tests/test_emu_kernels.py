"""Kernel index logic + host orchestration, executed by the CPU fiber emulator (tests/emu) and checked against
the reference's golden vectors.  These are logic tests of the kernel *sources*; the numerical parity tests
proper run on the GPU (test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2


@pytest.mark.parametrize("n", [32, 96])
def test_laplacian(emu_solver, gold, n):
    g = gold(f"operator_n{n}.npz")
    emu_solver.set_domain_size(n, source_location=[n // 3, n // 2])
    assert rel_l2(emu_solver.Lap(torch.tensor(g["u"])), g["Lu"]) < 1e-6
    assert rel_l2(emu_solver.source, g["source"]) < 1e-6       # hn_point_sources writes the exact map (the reference carries 4e-7 of FFT noise)


def test_laplacian_generic_radix(emu_solver):
    """N = 80 = 16 * 5 exercises the generic-radix butterfly."""
    from oracle import helmnet_oracle as O
    emu_solver.set_domain_size(80, source_location=[3, 4])
    u = torch.randn(1, 80, 80, 2, generator=torch.Generator().manual_seed(3))
    assert rel_l2(emu_solver.Lap(u), O.laplacian(u, O.make_operator(80, 8, 2.0, 1.0))) < 1e-6


def test_residual_n256_per_sample_sources(emu_solver):
    """The N = 256 residual kernels with per-sample point sources (the zero-source tile skip reads 8 column flags per tile)
    against the oracle."""
    from oracle import helmnet_oracle as O
    emu_solver.set_domain_size(256, source_location=[30, 128])
    locs = [[30, 128], [200, 5]]
    emu_solver.set_multiple_sources(locs)
    gen = torch.Generator().manual_seed(21)
    wf = torch.randn(2, 2, 256, 256, generator=gen)
    k_sq = 1.0 + torch.rand(2, 1, 256, 256, generator=gen)
    ref = O.get_residual(wf, k_sq, O.point_sources(256, locs), O.make_operator(256, 8, 2.0, 1.0))
    assert rel_l2(emu_solver.get_residual(wf, k_sq), ref) < 1e-6


@pytest.mark.parametrize("pml", [8, 12, 0])
def test_laplacian_n256_register_fft(emu_solver, pml):
    """N = 256 takes the register-resident 16 x 16 kernels (spectral256.cuh); pml 12 has threads owning two strip samples."""
    from oracle import helmnet_oracle as O
    old = emu_solver.hparams.PMLsize
    emu_solver.hparams.PMLsize = pml
    try:
        emu_solver.set_domain_size(256, source_location=[30, 128])
        u = torch.randn(1, 256, 256, 2, generator=torch.Generator().manual_seed(5))
        ref = O.laplacian(u, O.make_operator(256, pml, 2.0, 1.0)) if pml > 0 else None
        out = emu_solver.Lap(u)
        if pml > 0:
            assert rel_l2(out, ref) < 1e-6
        else:   # no PML: plain spectral Laplacian
            k = torch.tensor(O.wavenumbers(256)).float()
            uf = torch.fft.fftn(torch.view_as_complex(u), dim=(-2, -1))
            lap = torch.view_as_real(torch.fft.ifftn(-(k[None, :, None] ** 2 + k[None, None, :] ** 2) * uf, dim=(-2, -1)))
            assert rel_l2(out, lap) < 1e-6
    finally:
        emu_solver.hparams.PMLsize = old


@pytest.mark.parametrize("pml", [8, 13])
def test_laplacian_and_residual_n512_register_fft(emu_solver, pml):
    """N = 512: one warp per line, two 256-point register transforms + a radix-2 butterfly across the half-warps
    (spectral512.cuh); pml 13 has threads owning strip samples of both parities."""
    from oracle import helmnet_oracle as O
    old = emu_solver.hparams.PMLsize
    emu_solver.hparams.PMLsize = pml
    try:
        emu_solver.set_domain_size(512, source_location=[450, 256])
        gen = torch.Generator().manual_seed(11)
        u = torch.randn(1, 512, 512, 2, generator=gen)
        op = O.make_operator(512, pml, 2.0, 1.0)
        assert rel_l2(emu_solver.Lap(u), O.laplacian(u, op)) < 1e-6
        # residual r = L u + k_sq u - source through the fused column epilogue
        wf = torch.randn(1, 2, 512, 512, generator=gen)
        k_sq = 1.0 + torch.rand(1, 1, 512, 512, generator=gen)
        lu = O.laplacian(wf.permute(0, 2, 3, 1).contiguous(), op).permute(0, 3, 1, 2)
        ref = lu + k_sq * wf - emu_solver.source
        assert rel_l2(emu_solver.get_residual(wf, k_sq), ref) < 1e-6
    finally:
        emu_solver.hparams.PMLsize = old


def test_laplacian_and_residual_n1024_register_fft(emu_solver):
    """N = 1024: two warps per line, radix-2 across the warps around the 512-point warp transforms (spectral1024.cuh)."""
    from oracle import helmnet_oracle as O
    emu_solver.set_domain_size(1024, source_location=[60, 512])
    gen = torch.Generator().manual_seed(13)
    u = torch.randn(1, 1024, 1024, 2, generator=gen)
    op = O.make_operator(1024, 8, 2.0, 1.0)
    assert rel_l2(emu_solver.Lap(u), O.laplacian(u, op)) < 1e-6
    wf = torch.randn(1, 2, 1024, 1024, generator=gen)
    k_sq = 1.0 + torch.rand(1, 1, 1024, 1024, generator=gen)
    lu = O.laplacian(wf.permute(0, 2, 3, 1).contiguous(), op).permute(0, 3, 1, 2)
    assert rel_l2(emu_solver.get_residual(wf, k_sq), lu + k_sq * wf - emu_solver.source) < 1e-6


def test_unet_and_single_step(emu_solver, gold):
    g = gold("unet_step_n32.npz")
    s = emu_solver
    s.set_domain_size(32, source_location=[10, 16])
    s.f.set_states(torch.tensor(g["states_flat"]), flatten=True)
    d = s.f(torch.tensor(g["inp"]))
    assert rel_l2(d, g["d_wf"]) < 5e-6
    assert rel_l2(s.f.get_states(flatten=True), g["states_flat_out"]) < 5e-6
    s.f.set_states(torch.tensor(g["states_flat"]), flatten=True)
    up, res = s.single_step(torch.tensor(g["wf"]), torch.tensor(g["k_sq"]), torch.tensor(g["res"]))
    assert rel_l2(up, g["up_wf"]) < 1e-6 and rel_l2(res, g["new_res"]) < 1e-5
    assert rel_l2(s.get_residual(torch.tensor(g["wf"]), torch.tensor(g["k_sq"])), g["residual_of_wf"]) < 1e-6


def test_forward_per_sample_sources(emu_solver, gold):
    g = gold("traj_srcmap_n64.npz")
    s = emu_solver
    s.set_domain_size(64, source_map=torch.tensor(g["source"]))
    out = s.forward(torch.tensor(g["sos"]), num_iterations=6, return_wavefields=True, return_states=True)
    assert out["last_iteration"] == 5 and len(out["wavefields"]) == 6 and len(out["residuals"]) == 6 and len(out["states"]) == 6
    assert rel_l2(out["residual_rmse"], g["rmse"][:6]) < 1e-5
    rm = torch.stack([s.test_loss_function(r) for r in out["residuals"]])
    assert rel_l2(rm, out["residual_rmse"]) < 1e-6
    # n_steps continues a solve exactly where forward left it
    k_sq, _ = s.get_initials(torch.tensor(g["sos"]))
    s.f.set_states(out["states"][2], flatten=True)
    cont = s.n_steps(out["wavefields"][2], k_sq, out["residuals"][2], 3)
    assert rel_l2(cont["wavefields"][0], out["wavefields"][5]) < 1e-6
    assert rel_l2(s.f.get_states(flatten=True), out["states"][5]) < 1e-6


def test_forward_variable_src(emu_solver, gold):
    g = gold("traj_srcmap_n64.npz")
    s = emu_solver
    src = torch.tensor(g["source"])
    s.set_domain_size(64, source_map=src)
    sos = torch.tensor(g["sos"])
    a = s.forward_variable_src(sos, {"iteration": [2], "src_maps": [2 * src]}, num_iterations=4)
    # same thing by hand with the public pieces
    s.set_source_maps(src)
    o = s.forward(sos, num_iterations=2, return_states=True)
    s.set_source_maps(2 * src)
    k_sq, _ = s.get_initials(sos)
    res = s.get_residual(o["wavefields"][0], k_sq)
    b = s.n_steps(o["wavefields"][0], k_sq, res, 2)
    assert rel_l2(a["wavefields"][0], b["wavefields"][0]) < 1e-6
    assert a["residual_rmse"].shape == (4, 3)


def test_evaluation_driver_outputs(tmp_path, gold):
    """evaluate.py flow (get_model -> test_step -> test_epoch_end) writes the two result files with the reference's shapes."""
    import numpy as np
    from conftest import CKPT
    from emu_backend import EmuLib
    from helmnet_b200 import IterativeSolver, evaluate
    model = evaluate.get_model(CKPT, domain_size=32, source_location=[10, 16])
    model._backend = EmuLib()
    assert model.hparams.domain_size == 32 and model.source.shape == (1, 2, 32, 32)
    sos = torch.ones(3, 1, 32, 32)
    sos[:, :, 10:20, 5:25] = 1.4
    losses = evaluate.results_on_test_set(model, sos, batch_size=2, out_dir=str(tmp_path), max_iterations=4)
    a = np.load(tmp_path / "evolution_of_model_RMSE_on_test_set.npy")
    w = np.load(tmp_path / "evolution_of_wavefields_on_test_set.npy")
    assert a.shape == (3, 4) and w.shape == (3, 4, 2, 32, 32) and np.array_equal(a, losses)
    # same numbers as a direct forward
    ref = model.forward(sos, num_iterations=4, return_wavefields=True)
    assert rel_l2(torch.tensor(w[:, 3]), ref["wavefields"][3]) < 1e-6


@pytest.mark.parametrize("batch,H,pad,grid", [(32, 256, 8, 148), (256, 256, 8, 148), (37, 128, 8, 296), (256, 128, 6, 296), (9, 512, 6, 296),
                                              (100000, 256, 8, 148), (5, 64, 6, 296), (64, 8 * 512, 6, 296), (86, 32, 6, 296)])
def test_balanced_strip_partition(batch, H, pad, grid):
    """common.cuh: balanced_strip (the strip walk of the persistent tcgen05 kernels; executed here through a probe kernel on the
    emulator).  Every row of every image belongs to exactly one strip, strips start and end on even rows, a CTA's strips are in
    ascending order, and every CTA carries the same rows + pad * strips to within one strip start."""
    import ctypes as C
    from emu_backend import build_emu
    dll = C.CDLL(build_emu())
    cap = 3 + (batch * (H + pad) // grid) // (H + pad) + 3
    out = (C.c_int * (grid * cap * 3))()
    dll.emu_balanced_strips(batch, H, pad, grid, out, cap)
    a = np.frombuffer(out, dtype=np.int32).reshape(grid, cap, 3)
    covered = np.zeros((batch, H), np.int32) if batch * H < 5_000_000 else None
    rows_total, costs, last = 0, [], (-1, -1)
    for c in range(grid):
        cost = 0
        for i in range(cap):
            b, y0, R = (int(v) for v in a[c, i])
            if b < 0:
                break
            assert 0 <= b < batch and 0 <= y0 and R >= 2 and y0 + R <= H and y0 % 2 == 0 and R % 2 == 0
            assert (b, y0) > last
            last = (b, y0)
            if covered is not None:
                covered[b, y0:y0 + R] += 1
            rows_total += R
            cost += R + pad
        else:
            raise AssertionError("strip list not terminated")
        costs.append(cost)
    assert rows_total == batch * H
    if covered is not None:
        assert covered.min() == 1 and covered.max() == 1
    busy = [x for x in costs if x > 0]
    assert max(busy) - min(busy) <= 2 * pad + 4, (min(busy), max(busy))
