"""helmnet_b200/training.py (replay buffer + training_step around the differentiable n_steps) against the UNMODIFIED
reference's own ReplayBuffer / training_step (helmnet/hybridnet.py:385-505) run on CPU in the build container; the kernels of
this build are executed by the emulator.  Skipped where the reference tree is absent (the GPU box)."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from conftest import CKPT, ROOT, rel_l2

REF = os.environ.get("HELMNET_REFERENCE", "/root/reference")
N, BATCH, CAP, STEPS = 32, 2, 4, 2


def reference_solver():
    if not os.path.isdir(os.path.join(REF, "helmnet")):
        pytest.skip("reference tree not present")
    for p in (REF, os.path.join(ROOT, "oracle", "_stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from helmnet import IterativeSolver as RefSolver
    from helmnet.replaybuffer import Experience as RefExperience, ReplayBuffer as RefBuffer
    s = RefSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    s.hparams.source_location = [10, 16]
    s.set_domain_size(N, source_location=[10, 16])
    s.hparams.batch_size, s.hparams.buffer_size, s.hparams.unrolling_steps = BATCH, CAP, STEPS
    s.replaybuffer = RefBuffer(CAP)

    class Sink:
        def __getattr__(self, name):
            return lambda *a, **k: None
    s.logger = types.SimpleNamespace(experiment=Sink())
    s.trainer = types.SimpleNamespace(global_step=0, current_epoch=1)
    s.current_epoch = 1
    return s, RefExperience


def our_solver():
    from emu_backend import EmuLib
    from helmnet_b200 import IterativeSolver
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None, _backend=EmuLib())
    s.train()
    s.hparams.source_location = [10, 16]
    s.set_domain_size(N, source_location=[10, 16])
    s.hparams.batch_size, s.hparams.buffer_size, s.hparams.unrolling_steps = BATCH, CAP, STEPS
    return s


def training_maps():
    g = torch.Generator().manual_seed(3)
    sos = torch.ones(CAP, 1, N, N)
    sos[:, :, 8:20, 6:26] += 0.5 * torch.rand(CAP, 1, 12, 20, generator=g)
    return sos


def test_training_step_matches_the_reference():
    from helmnet_b200 import training as T
    ref, RefExperience = reference_solver()
    ours = our_solver()
    sos = training_maps()
    # fill both buffers the way train_dataloader does (hybridnet.py:196-218)
    with torch.no_grad():
        for i in range(CAP):
            ref.reset_source()
            k_sq, wf = ref.get_initials(sos[i:i + 1])
            ref.f.clear_states(wf)
            h = ref.f.get_states(flatten=True)
            res = ref.get_residual(wf, k_sq)
            ref.replaybuffer.append(RefExperience(wf[0], h[0], k_sq[0], res[0], ref.source[0], i * 10), i)
    buf = T.ReplayBuffer(CAP)
    T.fill_replay_buffer(ours, buf, sos)
    for i in range(CAP):
        e = ref.replaybuffer.buffer[i]
        assert buf.iterations[i] == e.iteration
        assert rel_l2(buf._store["residual"][i], e.residual) < 1e-6 and torch.equal(buf._store["k_sq"][i], e.k_sq)
    # one training step each, same RNG streams.  current_epoch = 1: maxiter = 21, so the slots that started at 20 / 30 restart
    np.random.seed(5); random.seed(5)
    ref_out = ref.training_step(sos[:BATCH], 0)
    ref_out["loss"].backward()
    np.random.seed(5); random.seed(5)
    out = T.training_step(ours, sos[:BATCH], buf, current_epoch=1)
    out["loss"].backward()
    assert abs(float(out["loss"]) - float(ref_out["loss"])) / abs(float(ref_out["loss"])) < 1e-5
    assert out["maxiter"] == 21 and 0 <= out["new_sos"] <= BATCH
    for i in range(CAP):
        e = ref.replaybuffer.buffer[i]
        assert buf.iterations[i] == e.iteration, i
        for f in ("wavefield", "hidden_state", "k_sq", "residual", "source"):
            a, b = buf._store[f][i], getattr(e, f)
            assert (float(b.norm()) == 0 and float(a.norm()) == 0) or rel_l2(a, b) < 1e-5, (i, f)
    g_ref = dict(ref.f.named_parameters())
    worst = max(rel_l2(p.grad, g_ref[k].grad) for k, p in ours.f.named_parameters())
    assert worst < 1e-3, worst       # the reference's own fp32 autograd is the yardstick here (slope gradients: 1e-4, test_emu_train)
    # clipping + optimizer step as the reference configures them
    T.on_after_backward(ours)
    opt, sched = T.configure_optimizers(ours)
    before = ours.f.inc.double_conv[0].weight.detach().clone()
    opt.step()
    assert not torch.equal(before, ours.f.inc.double_conv[0].weight.detach())
    assert isinstance(opt, torch.optim.Adam) and opt.defaults["betas"] == (0.9, 0.95)


def test_fit_runs_and_refreshes_the_buffer():
    from helmnet_b200 import training as T
    ours = our_solver()
    np.random.seed(1); random.seed(1); torch.manual_seed(1)
    hist = T.fit(ours, training_maps(), epochs=2)
    assert len(hist) == 2 and all(np.isfinite(h) for h in hist)
    with pytest.raises(RuntimeError):
        T.ReplayBuffer(3).sample(2)
