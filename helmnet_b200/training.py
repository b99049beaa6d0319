"""Training-side host logic around the differentiable ``n_steps``.

What the reference keeps inside its LightningModule (helmnet/hybridnet.py:172-176, 192-218, 250-284, 385-505) and its replay
buffer (helmnet/replaybuffer.py), without Lightning: the experience replay, one ``training_step`` (sample states from the
buffer, unroll ``unrolling_steps`` solver steps under autograd, loss, refresh the buffer), gradient clipping, the optimizer
and a small ``fit`` loop that stands in for the Trainer.  All arithmetic of the unroll and of its backward runs in the CUDA
kernels behind ``IterativeSolver.n_steps``; this file only moves tensors between the buffer and the solver.

The buffer is device-resident: one preallocated tensor per field ``[capacity, ...]``; sampling is one ``index_select`` per
field instead of stacking ``batch_size`` separately stored tensors, writing back one ``index_copy_``.  It keeps the reference's
interface (``append(experience, index)``, ``sample(batch_size)`` -> the same 7-tuple) and consumes NumPy's global RNG exactly
as the reference does, so a seeded run visits the same samples.
"""
from __future__ import annotations

import collections
import random
from typing import Optional

import numpy as np
import torch

Experience = collections.namedtuple("Experience", ["wavefield", "hidden_state", "k_sq", "residual", "source", "iteration"])

_FIELDS = ("wavefield", "hidden_state", "k_sq", "residual", "source")


class ReplayBuffer:
    """helmnet/replaybuffer.py:20-47 with tensor storage."""

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self._store = None                      # field -> [capacity, ...] tensor, allocated at the first write
        self.iterations = [None] * self.capacity

    def __len__(self):
        return self.capacity

    def _ensure(self, example: Experience):
        if self._store is None:
            self._store = {f: torch.empty((self.capacity,) + tuple(getattr(example, f).shape), dtype=getattr(example, f).dtype,
                                          device=getattr(example, f).device) for f in _FIELDS}

    def append(self, experience: Experience, index: int):
        self._ensure(experience)
        for f in _FIELDS:
            self._store[f][index].copy_(getattr(experience, f).detach())
        self.iterations[index] = int(experience.iteration)

    def write(self, indices, wavefield, hidden_state, k_sq, residual, source, iterations):
        """Batched ``append``: row j of every tensor goes to slot ``indices[j]``."""
        self._ensure(Experience(wavefield[0], hidden_state[0], k_sq[0], residual[0], source[0], 0))
        idx = torch.as_tensor(np.asarray(indices), dtype=torch.long, device=self._store["wavefield"].device)
        for f, t in zip(_FIELDS, (wavefield, hidden_state, k_sq, residual, source)):
            self._store[f].index_copy_(0, idx, t.detach().to(self._store[f].dtype))
        for i, it in zip(indices, iterations):
            self.iterations[int(i)] = int(it)

    def sample(self, batch_size: int):
        if self._store is None or any(it is None for it in self.iterations):
            raise RuntimeError("the replay buffer must be filled before sampling (fill_replay_buffer)")
        batch_size = min(int(batch_size), self.capacity)
        indices = np.random.choice(self.capacity, batch_size, replace=False)
        idx = torch.as_tensor(indices, dtype=torch.long, device=self._store["wavefield"].device)
        out = [self._store[f].index_select(0, idx) for f in _FIELDS]
        return (*out, tuple(self.iterations[int(i)] for i in indices), indices)


def _initial_state(solver, sos):
    """hybridnet.py:201-208 / 454-458 for a batch: zero wavefield and hidden states, k_sq, residual of the zero field."""
    with torch.no_grad():
        k_sq, wf = solver.get_initials(sos)
        solver.f.clear_states(wf)
        h = solver.f.get_states(flatten=True)
        res = solver.get_residual(wf, k_sq)
    return wf, h, k_sq, res


def _default_source(solver):
    """``reset_source()`` (hybridnet.py:155-159) leaves a batch of sampled source maps in place when the source module already sits
    at the default location; the rest state of a restarted sample is defined against ONE map (the reference broadcasts and
    takes row 0, which is the same map whenever every buffer entry carries the default source, as in its training loop)."""
    solver.reset_source()
    if solver.source.shape[0] != 1:
        with torch.no_grad():
            solver.set_source()


def fill_replay_buffer(solver, buffer: ReplayBuffer, sos_train: torch.Tensor, chunk: int = 64):
    """hybridnet.py:196-218: slot i starts from the rest state of training map i with the iteration count 10 i."""
    if sos_train.shape[0] < len(buffer):
        raise ValueError(f"{len(buffer)} buffer slots but only {sos_train.shape[0]} training maps")
    _default_source(solver)
    for lo in range(0, len(buffer), chunk):
        idx = list(range(lo, min(lo + chunk, len(buffer))))
        sos = sos_train[idx].to(solver.device).type_as(solver.source)
        wf, h, k_sq, res = _initial_state(solver, sos)
        src = solver.source.detach()[:1].expand(len(idx), -1, -1, -1)
        buffer.write(idx, wf, h, k_sq, res, src, [10 * i for i in idx])


def training_step(solver, sos_batch: torch.Tensor, buffer: ReplayBuffer, current_epoch: int = 0):
    """hybridnet.py:385-505 without the logging: returns ``{"loss", "rel_loss", "maxiter", "new_sos", "indices"}``; the caller
    runs ``loss.backward()`` (then ``on_after_backward`` and the optimizer step)."""
    hp = solver.hparams
    maxiter = min(current_epoch * 20 + 1, hp.max_iterations)
    wavefields, h_states, k_sqs, residual, sources, timesteps, indices = buffer.sample(hp.batch_size)
    solver.set_source_maps(sources)
    solver.f.set_states(h_states, flatten=True)
    out = solver.n_steps(wavefields, k_sqs, residual, hp.unrolling_steps, True, True)
    loss_f = torch.cat(out["residuals"]).pow(2)
    loss = 1e4 * loss_f.mean()
    rel_loss = loss_f.detach().mean((1, 2, 3)).sqrt().mean()

    # refresh the buffer: the state after a random one of the unrolled steps goes back into the sample's slot unless its residual
    # blew up or it ran out of iterations -- then the slot restarts from the rest state of a random map of this batch
    n = len(indices)
    iteration = int(np.random.choice(len(out["residuals"])))
    wf_k, h_k, res_k = (out[key][iteration].detach() for key in ("wavefields", "states", "residuals"))
    new_t = [int(t) + iteration + 1 for t in timesteps]
    alive = (res_k.pow(2).mean((1, 2, 3)) < 1).tolist()
    keep = [j for j in range(n) if alive[j] and new_t[j] < maxiter]
    drop = [j for j in range(n) if not (alive[j] and new_t[j] < maxiter)]
    if keep:
        kj = torch.as_tensor(keep, dtype=torch.long, device=wf_k.device)
        buffer.write([indices[j] for j in keep], wf_k.index_select(0, kj), h_k.index_select(0, kj), k_sqs.index_select(0, kj),
                     res_k.index_select(0, kj), sources.index_select(0, kj), [new_t[j] for j in keep])
    if drop:
        _default_source(solver)
        picks = torch.stack([random.choice(sos_batch) for _ in drop]).to(solver.device).type_as(solver.source)
        wf0, h0, k0, r0 = _initial_state(solver, picks)
        buffer.write([indices[j] for j in drop], wf0, h0, k0, r0, solver.source.detach()[:1].expand(len(drop), -1, -1, -1), [0] * len(drop))
    return {"loss": loss, "rel_loss": rel_loss, "maxiter": maxiter, "new_sos": len(drop), "indices": indices, "iteration": iteration}


def on_after_backward(solver):
    """hybridnet.py:172-176."""
    if solver.hparams.gradient_clip_val > 0:
        torch.nn.utils.clip_grad_value_(solver.parameters(), solver.hparams.gradient_clip_val)


def configure_optimizers(solver):
    """hybridnet.py:250-284: Adam(betas = (0.9, 0.95)) + ReduceLROnPlateau on the epoch mean of the training loss."""
    hp = solver.hparams
    if str(hp.optimizer).lower() != "adam":
        raise NotImplementedError("The optimizer {} is not implemented".format(hp.optimizer))
    if hp.minimum_learning_rate > hp.learning_rate:
        raise ValueError("Minimum learning rate ({}) must be smaller than the starting learning rate ({})".format(
            hp.minimum_learning_rate, hp.learning_rate))
    params = [p for p in solver.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=hp.learning_rate, betas=(0.9, 0.95), weight_decay=hp.weight_decay)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=10, min_lr=hp.minimum_learning_rate)
    return opt, sched


def fit(solver, sos_train: torch.Tensor, epochs: int = 1, buffer: Optional[ReplayBuffer] = None, steps_per_epoch: Optional[int] = None):
    """A minimal stand-in for ``Trainer.fit`` (train.py:49-71): fills the buffer, then per epoch walks the training maps in
    batches (drop_last), one optimizer step per batch, scheduler step on the epoch's mean loss.  Returns the per-epoch losses."""
    hp = solver.hparams
    solver.train()
    buffer = buffer if buffer is not None else ReplayBuffer(hp.buffer_size)
    fill_replay_buffer(solver, buffer, sos_train)
    opt, sched = configure_optimizers(solver)
    history = []
    for epoch in range(epochs):
        losses = []
        nb = sos_train.shape[0] // hp.batch_size
        for b in range(nb if steps_per_epoch is None else min(nb, steps_per_epoch)):
            opt.zero_grad(set_to_none=True)
            out = training_step(solver, sos_train[b * hp.batch_size:(b + 1) * hp.batch_size], buffer, epoch)
            out["loss"].backward()
            on_after_backward(solver)
            opt.step()
            losses.append(out["loss"].detach())
        mean = torch.stack(losses).mean()
        sched.step(mean)
        history.append(float(mean))
    return history
