import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")
CKPT = os.path.join(GOLD, "jcp_paper_trained_weights_slim.ckpt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def record(name, **vals):
    """Append measured parity numbers to gpurun_out/parity_measured.jsonl (copied into profiles/ per round)."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if not isinstance(v, (list, str, int)) else v) for k, v in vals.items()}}) + "\n")
    except OSError:
        pass


@pytest.fixture(scope="session")
def gold():
    def load(name):
        return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}
    return load


@pytest.fixture(scope="session")
def f_weights():
    from helmnet_b200.checkpoint import load_checkpoint
    sd = load_checkpoint(CKPT)["state_dict"]
    return {k[2:]: v for k, v in sd.items() if k.startswith("f.")}


@pytest.fixture(scope="session")
def _cuda_solver_base():
    from helmnet_b200 import IterativeSolver
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    s.freeze()
    s.to("cuda:0")
    return s


@pytest.fixture(params=[0, 1, 2], ids=["simt", "tcgen05", "tcgen05-fused"])
def cuda_solver(request, _cuda_solver_base):
    """IterativeSolver on cuda:0 through the product library (fails loudly if it is missing), once per
    convolution engine: fp32 CUDA cores, tcgen05 split-fp16 (one kernel per conv) and tcgen05 with fused DoubleConvs."""
    _cuda_solver_base.set_engine(request.param)
    yield _cuda_solver_base
    _cuda_solver_base.sync_check()


@pytest.fixture(scope="session")
def emu_solver():
    """Same host code, kernels executed by the CPU fiber emulator (tests/emu) -- logic tests only."""
    from emu_backend import EmuLib
    from helmnet_b200 import IterativeSolver
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None, _backend=EmuLib())
    s.freeze()
    return s
