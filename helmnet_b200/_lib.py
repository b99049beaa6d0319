"""ctypes binding of libhelmnet_sm100.so (include/helmnet_sm100.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m helmnet_b200.build``.  If it is
missing, or no Blackwell GPU is present, everything here raises: there is no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HN_NUM_WEIGHTS = 48160
_HERE = os.path.dirname(os.path.abspath(__file__))


class LibraryMissingError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("HELMNET_SM100_LIB", os.path.join(_HERE, "csrc", "libhelmnet_sm100.so"))


_P = C.c_void_p
_SIGS = {
    "hn_version": (C.c_char_p, []),
    "hn_last_error": (C.c_char_p, []),
    "hn_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "hn_destroy": (C.c_int, [_P]),
    "hn_load_weights": (C.c_int, [_P, _P, C.c_size_t]),
    "hn_set_source": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int64), _P]),
    "hn_point_sources": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_double, C.c_double, C.c_int, _P, _P]),
    "hn_reset": (C.c_int, [_P, _P, C.c_int, _P]),
    "hn_set_state": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    "hn_run": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    "hn_step_backward": (C.c_int, [_P] + [_P] * 11 + [C.c_int, _P]),
    "hn_get": (C.c_int, [_P, _P, _P, _P, _P]),
    "hn_get_states": (C.c_int, [_P, _P, C.c_int, _P]),
    "hn_residual": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    "hn_laplacian": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "hn_unet": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "hn_state_len": (C.c_int, [_P]),
    "hn_launch_count": (C.c_int64, [_P]),
    "hn_kernels_per_iteration": (C.c_int, [_P]),
    "hn_debug_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int, _P]),
    "hn_set_engine": (C.c_int, [_P, C.c_int]),
    "hn_sync_check": (C.c_int, [_P, _P]),
    "hn_profile_layer": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), _P]),
    "hn_profile_iteration": (C.c_int, [_P, C.POINTER(C.c_float), _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


class HelmnetError(RuntimeError):
    pass


class HelmnetLib:
    """Thin typed wrapper; ``check(rc)`` turns a negative status into an exception carrying hn_last_error()."""

    requires_cuda = True

    def __init__(self, path: Optional[str] = None):
        path = path or lib_path()
        if not os.path.exists(path):
            raise LibraryMissingError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). helmnet_b200 has no CPU or PyTorch fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def last_error(self) -> str:
        return (self.hn_last_error() or b"").decode()

    def check(self, rc: int, what: str = "") -> int:
        if rc < 0:
            raise HelmnetError(f"{what or 'libhelmnet_sm100'} failed ({rc}): {self.last_error()}")
        return rc

    def version(self) -> str:
        return self.hn_version().decode()

    # tensors must live where the kernels can reach them
    def check_tensor(self, t, name="tensor"):
        if not t.is_cuda:
            raise HelmnetError(f"{name} must be a CUDA tensor: helmnet_b200 has no CPU path")

    def stream_for(self, device) -> int:
        import torch
        return torch.cuda.current_stream(device).cuda_stream


_default: Optional[HelmnetLib] = None


def default_lib() -> HelmnetLib:
    global _default
    if _default is None:
        _default = HelmnetLib()
    return _default
