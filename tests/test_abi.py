"""C-ABI surface: the built library exports every symbol include/helmnet_sm100.h declares, and the product
path refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from conftest import CKPT, ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "helmnet_sm100.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hn_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_header_symbols():
    from helmnet_b200 import build
    from helmnet_b200._lib import EXPORTED_SYMBOLS
    path = build.build()
    dll = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(dll, s), s
    assert set(EXPORTED_SYMBOLS) == set(syms)
    dll.hn_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.hn_version()


def test_sass_is_sm100_and_uses_ffma2():
    import subprocess
    from helmnet_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out or "sm_100" in out
    assert out.count("FFMA2") > 1000


def test_sass_census_of_the_tensor_core_path():
    """The default engine is tcgen05/TMEM/TMA code, not a recompiled mma.sync path: UTCHMMA (tcgen05.mma kind::f16),
    LDTM/STTM (tcgen05.ld/st), UBLKCP (cp.async.bulk), ACQBULK (griddepcontrol.wait) per kernel family."""
    import subprocess
    from helmnet_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    census, name = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            census[name] = {"UTCHMMA": 0, "LDTM": 0, "STTM": 0, "UBLKCP": 0, "ACQBULK": 0, "UTCBAR": 0}
        elif name:
            for k_ in census[name]:
                if k_ in line:
                    census[name][k_] += 1
    fam = lambda key: [v for k_, v in census.items() if key in k_]
    for key in ("dconv_tcf_kernel", "down_tcr_kernel", "up_tcr_kernel", "conv3x3_tcr_kernel"):
        ks = fam(key)
        assert ks, key
        for v in ks:
            assert v["UTCHMMA"] > 0 and v["LDTM"] > 0 and v["UBLKCP"] > 0 and v["ACQBULK"] > 0, (key, v)
    assert sum(v["UTCHMMA"] for v in census.values()) > 500
    assert not any("HMMA." in l and "UTCHMMA" not in l for l in out.splitlines()), "legacy mma.sync found"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from helmnet_b200 import IterativeSolver
    from helmnet_b200._lib import HelmnetError, HelmnetLib
    lib = HelmnetLib()
    ctx = ctypes.c_void_p()
    rc = lib.hn_create(ctypes.byref(ctx), 0, 64, 1, 8, 2.0, 1.0, 1.0)
    assert rc == -2 and "no CPU fallback" in lib.last_error()
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    s.set_domain_size(64, source_location=[10, 10])
    with pytest.raises(HelmnetError):
        s.forward(torch.ones(1, 1, 64, 64), num_iterations=1)


def test_bad_arguments_report_errors():
    from emu_backend import EmuLib
    lib = EmuLib()
    ctx = ctypes.c_void_p()
    assert lib.hn_create(ctypes.byref(ctx), 0, 50, 1, 8, 2.0, 1.0, 1.0) == -1 and "multiple of 16" in lib.last_error()
    assert lib.hn_create(ctypes.byref(ctx), 0, 32, 0, 8, 2.0, 1.0, 1.0) == -1
    assert lib.hn_create(ctypes.byref(ctx), 0, 32, 2, 8, 2.0, 1.0, 1.0) == 0
    assert lib.hn_run(ctx, 1, None, None, None, None, None) == -3          # no solve state yet
    blob = (ctypes.c_float * 10)()
    assert lib.hn_load_weights(ctx, blob, 10) == -1
    assert lib.hn_reset(ctx, blob, 1, None) == -3                            # weights / source missing
    assert lib.hn_state_len(ctx) == 32 * 32 + 16 * 16 + 8 * 8 + 4 * 4
    assert lib.hn_destroy(ctx) == 0
