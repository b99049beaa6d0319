"""Host-side drop-in surface (no kernels): checkpoint loading without Lightning, hparams, state packing."""
import os

import numpy as np
import pytest
import torch

from conftest import CKPT, GOLD

REF_CKPT = "/root/reference/trained_models/jcp_paper_trained_weights.ckpt"


def test_slim_checkpoint_loads_and_matches_hparams():
    import json
    from helmnet_b200 import IterativeSolver
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    hp = json.load(open(os.path.join(GOLD, "hparams.json")))
    for k in ("domain_size", "k", "omega", "PMLsize", "sigma_max", "max_iterations", "source_amplitude", "features", "depth"):
        assert s.hparams[k] == hp[k]
    assert s.hparams.test_data_path is None and s.hparams.architecture == "custom_unet"
    assert s.f.weight_blob().numel() == 48160
    assert s.source.shape == (1, 2, 96, 96) and float(s.source[0, 0, 82, 48]) == 10.0
    assert s.sigmas.shape == (2, 96, 96) and float(s.sigmas[0, 0, 0]) == 2.0 and float(s.sigmas[1, 0, 5]) == 2.0


@pytest.mark.skipif(not os.path.exists(REF_CKPT), reason="reference checkpoint not mounted (GPU box)")
def test_legacy_lightning_checkpoint_loads_identically():
    from helmnet_b200 import IterativeSolver
    a = IterativeSolver.load_from_checkpoint(REF_CKPT, strict=False, test_data_path=None)
    b = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    assert torch.equal(a.f.weight_blob(), b.f.weight_blob()) and torch.equal(a.source, b.source)
    with pytest.raises(RuntimeError):   # strict=True must complain about the stale Lap.* keys like Lightning does
        IterativeSolver.load_from_checkpoint(REF_CKPT, strict=True, test_data_path=None)


def test_state_dict_names_match_reference_layout():
    from helmnet_b200 import IterativeSolver
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    keys = [k for k in s.state_dict() if k.startswith("f.")]
    ck = [k for k in torch.load(CKPT, weights_only=False)["state_dict"] if k.startswith("f.")]
    assert keys == ck


def test_state_packing_roundtrip():
    from helmnet_b200 import IterativeSolver
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    s.set_domain_size(64, source_location=[5, 6])
    assert s.f.states_dimension == [64, 32, 16, 8] and s.f.total_state_length == 5440
    assert [e.domain_size for e in s.f.enc] == [64, 32, 16, 8]
    flat = torch.randn(3, 2, 5440)
    s.f.set_states(flat, flatten=True)
    assert torch.equal(s.f.get_states(flatten=True), flat)
    s.f.clear_states(torch.zeros(3, 2, 64, 64))
    assert all(float(h.abs().sum()) == 0 for h in s.f.get_states())
    assert s.source.shape == (1, 2, 64, 64) and not s.source.is_contiguous()
    k_sq, wf = s.get_initials(torch.full((2, 1, 64, 64), 2.0))
    assert float(k_sq[0, 0, 0, 0]) == 0.25 and wf.shape == (2, 2, 64, 64)
    assert torch.allclose(s.test_loss_function(torch.ones(2, 2, 4, 4)), torch.ones(2))


def test_unsupported_architecture_is_refused():
    from helmnet_b200 import HybridNet
    with pytest.raises(NotImplementedError):
        HybridNet(features=16)
