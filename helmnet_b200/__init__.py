"""helmnet_b200 -- B200-native (sm_100a) implementation of helmnet's inference inner loop.

Drop-in for the reference's public class on that path::

    from helmnet_b200 import IterativeSolver
    solver = IterativeSolver.load_from_checkpoint("trained_models/jcp_paper_trained_weights.ckpt",
                                                  strict=False, test_data_path=None)
    solver.freeze(); solver.to("cuda:0")
    solver.set_domain_size(256, source_location=[30, 128])
    out = solver.forward(sos_maps, num_iterations=1000)

Everything numerical runs in hand-written CUDA behind the C ABI of
``helmnet_b200/csrc/libhelmnet_sm100.so`` (``include/helmnet_sm100.h``).  There is no CPU path: using the
solver without the built library or without a Blackwell GPU raises.
"""
from .checkpoint import HParams, load_checkpoint
from .solver import HybridNet, IterativeSolver
from .source import SourceModule
from .training import Experience, ReplayBuffer
from ._lib import HelmnetLib, LibraryMissingError, lib_path

__all__ = ["IterativeSolver", "HybridNet", "SourceModule", "HParams", "load_checkpoint", "HelmnetLib",
           "LibraryMissingError", "lib_path", "Experience", "ReplayBuffer"]
__version__ = "0.1.0"
