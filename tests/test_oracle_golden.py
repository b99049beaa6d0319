"""The oracle restatement vs golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import helmnet_oracle as O


@pytest.mark.parametrize("n", [32, 96])
def test_operator_tables_and_laplacian(gold, n):
    g = gold(f"operator_n{n}.npz")
    op = O.make_operator(n, 8, 2.0, 1.0)
    assert np.array_equal(op["ax"][0, 0].numpy(), g["ax"]) and np.array_equal(op["bx"][0, 0].numpy(), g["bx"])
    assert np.array_equal(op["ay"][0, :, 0].numpy(), g["ay"]) and np.array_equal(op["by"][0, :, 0].numpy(), g["by"])
    assert np.array_equal(op["kx"][0, 0, :, 1].numpy(), g["kx"]) and np.array_equal(op["kx_sq"][0, 0, :, 0].numpy(), g["kx_sq"])
    assert np.array_equal(op["sigmas"].numpy(), g["sigmas"])
    Lu = O.laplacian(torch.tensor(g["u"]), op)
    assert np.array_equal(Lu.numpy(), g["Lu"])          # same torch ops in the same order: bit exact
    src = O.point_source(n, [n // 3, n // 2])
    assert np.array_equal(src.numpy(), g["source"])


def test_unet_and_single_step(gold, f_weights):
    g = gold("unet_step_n32.npz")
    n = 32
    states = O.unflatten_states(torch.tensor(g["states_flat"]), n)
    d_wf, new_states = O.unet_forward(f_weights, torch.tensor(g["inp"]), states)
    assert rel_l2(d_wf, g["d_wf"]) < 1e-6
    assert rel_l2(O.flatten_states(new_states), g["states_flat_out"]) < 1e-6
    orc = O.Oracle(f_weights, n)
    orc.set_source(torch.tensor(g["source"]))
    up, res, _ = orc.single_step(torch.tensor(g["wf"]), torch.tensor(g["k_sq"]), torch.tensor(g["res"]), states)
    assert rel_l2(up, g["up_wf"]) < 1e-6 and rel_l2(res, g["new_res"]) < 1e-5
    assert rel_l2(orc.residual(torch.tensor(g["wf"]), torch.tensor(g["k_sq"])), g["residual_of_wf"]) < 1e-6


def test_trajectory_n96(gold, f_weights):
    g = gold("traj_n96_b2.npz")
    orc = O.Oracle(f_weights, 96)
    orc.set_source(O.point_source(96, [82, 48]))
    out = orc.forward(torch.tensor(g["sos"]), 40, keep_wavefields=True)
    assert rel_l2(out["rmse"], g["rmse"]) < 1e-5
    for i, k in enumerate(g["keep"]):
        assert rel_l2(out["wavefields"][k], g["wavefields"][i]) < 1e-5


def test_trajectory_bench_workload(gold, f_weights):
    """bench.py's workload (first two synthetic 256^2 maps, source [30,128]): the oracle against the unmodified reference,
    and the fixture against the generator bench.py uses."""
    from helmnet_b200.synthetic import synthetic_sos
    g = gold("traj_bench_n256_b2.npz")
    assert np.array_equal(synthetic_sos(32, 256, seed=1)[:2].numpy(), g["sos"])
    orc = O.Oracle(f_weights, 256)
    orc.set_source(O.point_source(256, [30, 128]))
    out = orc.forward(torch.tensor(g["sos"]), 12, keep_wavefields=True)
    assert rel_l2(out["rmse"], g["rmse"]) < 1e-5
    for i, k in enumerate(g["keep"]):
        assert rel_l2(out["wavefields"][k], g["wavefields"][i]) < 1e-5


def test_trajectory_source_maps(gold, f_weights):
    g = gold("traj_srcmap_n64.npz")
    orc = O.Oracle(f_weights, 64)
    orc.set_source(torch.tensor(g["source"]))
    out = orc.forward(torch.tensor(g["sos"]), 30)
    assert rel_l2(out["rmse"], g["rmse"]) < 1e-5 and rel_l2(out["wavefield"], g["wavefield"]) < 1e-5


def test_readme_landmarks(gold):
    """SURVEY.md 8c landmarks of the README lens run, measured on the reference."""
    g = gold("traj_readme_n256.npz")
    r = g["rmse"][:, 0]
    assert abs(r[0] - 6.1573e-3) < 2e-6 and abs(r[10] - 2.5914e-3) < 2e-6 and abs(r[50] - 1.0912e-3) < 2e-6
    assert int(np.argmax(r < 1e-3)) == 52


def test_fp64_arbiter_is_close(gold, f_weights):
    """fp32 vs fp64 oracle: the iteration is contractive, so the 1e-5 bar is meaningful."""
    g = gold("traj_srcmap_n64.npz")
    o32, o64 = O.Oracle(f_weights, 64), O.Oracle(f_weights, 64, dtype=torch.float64)
    for o in (o32, o64):
        o.set_source(torch.tensor(g["source"]))
    a = o32.forward(torch.tensor(g["sos"]), 10)["wavefield"]
    b = o64.forward(torch.tensor(g["sos"]), 10)["wavefield"]
    assert rel_l2(a, b) < 1e-5
