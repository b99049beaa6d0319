"""Drop-in ``IterativeSolver`` for helmnet's inference path, backed by libhelmnet_sm100.so.

API parity target: ``helmnet.IterativeSolver`` (reference helmnet/hybridnet.py) as used by README.md:58-83,
examples/simple_scattering.py, evaluate.py and support_functions.fig_generic:

    load_from_checkpoint / freeze / to / device / hparams / set_domain_size / forward / n_steps /
    single_step / get_residual / apply_laplacian / get_initials / test_loss_function / set_laplacian /
    setup_source / set_source_maps / set_source / reset_source / set_multiple_sources /
    forward_variable_src / f.{state_dict, init_by_size, get_states, set_states, clear_states, ...}

PyTorch is the host here (device memory, streams, parameter containers).  All arithmetic of the
iteration -- UNet update, spectral Laplacian/PML residual, residual norm -- is enqueued through the C ABI on
the caller's current CUDA stream.  No Lightning, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import inspect
import os
import weakref
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .checkpoint import HParams, load_checkpoint
from .source import SourceModule


# --------------------------------------------------------------------------------------------------
# parameter containers with the reference's state_dict names (helmnet/architectures.py:63-84, 186-252, 317-388)
# --------------------------------------------------------------------------------------------------
class DoubleConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.PReLU(),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
        )


class OutConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)


class EncoderBlock(nn.Module):
    """Weights of one encoder level + the Python-visible hidden state (``.state``)."""

    def __init__(self, num_features: int, state_size: int = 2, domain_size: int = 0):
        super().__init__()
        self.state_size = state_size
        self.num_features = num_features
        self.domain_size = domain_size
        self.use_state = True
        self.conv_signal = DoubleConv(num_features + state_size, num_features)
        self.down = nn.Conv2d(num_features, num_features, kernel_size=8, padding=3, stride=2)
        self.conv_state = DoubleConv(num_features + state_size, state_size)
        self.state: Optional[torch.Tensor] = None

    def set_state(self, state):
        self.state = state

    def get_state(self):
        return self.state

    def clear_state(self, x):
        self.state = torch.zeros([x.shape[0], 2, self.domain_size, self.domain_size], device=x.device)


class HybridNet(nn.Module):
    """The learned optimizer.  Holds the trained parameters; ``forward`` runs the CUDA UNet."""

    def __init__(self, activation_function: str = "prelu", depth: int = 4, domain_size: int = 96, features: int = 8,
                 inchannels: int = 6, state_channels: int = 2, state_depth: int = 4):
        super().__init__()
        if (activation_function.lower(), depth, features, inchannels, state_channels, state_depth) != ("prelu", 4, 8, 6, 2, 4):
            raise NotImplementedError(
                "libhelmnet_sm100 implements the shipped architecture only: prelu, depth=4, features=8, "
                "inchannels=6, state_channels=2, state_depth=4")
        self.activation_function = activation_function
        self.depth, self.domain_size, self.features = depth, domain_size, features
        self.inchannels, self.state_channels, self.state_depth = inchannels, state_channels, state_depth
        self.init_by_size()
        self.inc = DoubleConv(inchannels, features)
        self.enc = nn.ModuleList([EncoderBlock(features, state_channels, self.states_dimension[d]) for d in range(depth)])
        self.decode = nn.ModuleList([DoubleConv(features + features * (i < depth), features) for i in range(depth + 1)])
        self.up = nn.ModuleList([nn.ConvTranspose2d(features, features, kernel_size=8, padding=3, output_padding=0, stride=2)
                                 for _ in range(depth)])
        self.outc = OutConv(features, 2)
        self._owner = None  # weakref to the IterativeSolver that executes this net

    # -- state packing (reference architectures.py:390-437) ------------------------------------------------
    def init_by_size(self):
        self.states_dimension = [self.domain_size // 2 ** d for d in range(self.depth)]
        self.total_state_length = sum(s * s for s in self.states_dimension)
        self.state_boundaries, start = [], 0
        for s in self.states_dimension:
            self.state_boundaries.append([start, start + s * s])
            start += s * s

    def get_states(self, flatten: bool = False):
        h = [e.get_state() for e in self.enc]
        return self.flatten_state(h) if flatten else h

    def clear_states(self, x):
        for e in self.enc:
            e.clear_state(x)

    def set_states(self, states, flatten: bool = False):
        h = self.unflatten_state(states) if flatten else states
        for e, s in zip(self.enc[: len(h)], h):
            e.set_state(s)

    def flatten_state(self, h_list):
        return torch.cat([x.reshape(x.shape[0], x.shape[1], -1) for x in h_list], 2)

    def unflatten_state(self, h_flatten):
        b, c = h_flatten.shape[0], h_flatten.shape[1]
        return [h_flatten[:, :, lo:hi].reshape(b, c, s, s) for (lo, hi), s in zip(self.state_boundaries, self.states_dimension)]

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            owner.sync_weights()          # evaluate.py:62 copies weights with new_model.f.load_state_dict(...)
        return res

    def weight_blob(self) -> torch.Tensor:
        """The 48,160 parameters in state_dict order, the layout hn_load_weights expects."""
        return torch.cat([v.detach().reshape(-1).float().cpu() for v in self.state_dict().values()]).contiguous()

    def forward(self, x):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("HybridNet.forward needs its IterativeSolver (CUDA context owner)")
        return owner._unet_forward(x)


class SpectralLaplacian(nn.Module):
    """``solver.Lap``: callable [B,H,W,2] -> [B,H,W,2] (reference spectral.py:246-262) + ``sigmas()``."""

    def __init__(self, domain_size: int, PMLsize: int, k: float, sigma_max: float, owner=None):
        super().__init__()
        self.domain_size, self.PMLsize, self.k, self.sigma_max = domain_size, PMLsize, k, sigma_max
        coord = np.arange(PMLsize)
        prof = sigma_max * np.abs(1 - coord / PMLsize) ** 2 if PMLsize > 0 else np.zeros((0,))
        sigma = np.zeros((domain_size,))
        if PMLsize > 0:
            sigma[:PMLsize] = prof
            sigma[-PMLsize:] = prof[::-1]
        sig = torch.tensor(sigma).float()
        self.register_buffer("sigma_x", sig[None, :].repeat(domain_size, 1))   # varies along W
        self.register_buffer("sigma_y", sig[:, None].repeat(1, domain_size))   # varies along H
        self._owner = weakref.ref(owner) if owner is not None else None

    def sigmas(self):
        return self.sigma_x, self.sigma_y

    def forward(self, x):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("SpectralLaplacian needs its IterativeSolver (CUDA context owner)")
        return owner.apply_laplacian(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)



# --------------------------------------------------------------------------------------------------
# training unroll: one solver step as an autograd node (reference hybridnet.py:558-623 under autograd)
# --------------------------------------------------------------------------------------------------
class _SingleStepFn(torch.autograd.Function):
    """``IterativeSolver.single_step`` (hybridnet.py:558-584) as ONE node of the autograd graph.

    forward: the inference kernels (hn_set_state + hn_run(1) + hn_get).  backward: ``hn_step_backward`` -- the step is
    recomputed in fp32 with every pre-activation kept and differentiated by hand-written CUDA kernels (train.cuh): gradients
    of the wavefield, the residual and the flattened hidden state, and of the 88 parameter tensors of ``solver.f`` (passed as
    ``*params`` in state_dict order so that autograd routes their gradients).  k_sq and the source get no gradient, as in the
    reference's training_step (hybridnet.py:385-410), where neither requires one.
    """

    @staticmethod
    def forward(ctx, solver, wavefield, k_sq, residual, hflat, *params):
        wf, ks, rs, hf = (solver._prep(t, n) for t, n in ((wavefield, "wavefield"), (k_sq, "k_sq"), (residual, "residual"), (hflat, "state")))
        b, lib = wf.shape[0], solver.lib
        c = solver._ensure_ctx(b)
        lib.check(lib.hn_set_state(c, solver._ptr(wf), solver._ptr(rs), solver._ptr(ks), solver._ptr(hf), b, solver._stream()), "hn_set_state")
        lib.check(lib.hn_run(c, 1, solver._ptr(None), solver._ptr(None), solver._ptr(None), solver._ptr(None), solver._stream()), "hn_run")
        new_wf, new_res, new_h = torch.empty_like(wf), torch.empty_like(rs), torch.empty_like(hf)
        lib.check(lib.hn_get(c, solver._ptr(new_wf), solver._ptr(new_res), solver._ptr(new_h), solver._stream()), "hn_get")
        ctx.solver = solver
        ctx.domain_size = int(solver.hparams.domain_size)
        ctx.save_for_backward(wf, ks, rs, hf)
        ctx.param_shapes = [tuple(p.shape) for p in params]
        return new_wf, new_res, new_h

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_wf, g_res, g_h):
        solver = ctx.solver
        if int(solver.hparams.domain_size) != ctx.domain_size:
            raise RuntimeError(f"the solver's domain size changed from {ctx.domain_size} to {solver.hparams.domain_size} between "
                               "n_steps and backward(): the step cannot be differentiated on the new operator")
        wf, ks, rs, hf = ctx.saved_tensors
        b, lib = wf.shape[0], solver.lib
        c = solver._ensure_ctx(b)
        up = [None if g is None else solver._prep(g, "gradient") for g in (g_wf, g_res, g_h)]
        gwf_in, gres_in, gh_in = torch.empty_like(wf), torch.empty_like(rs), torch.empty_like(hf)
        gp = torch.zeros(_lib.HN_NUM_WEIGHTS, device=wf.device)
        ptr = solver._ptr
        lib.check(lib.hn_step_backward(c, ptr(wf), ptr(rs), ptr(ks), ptr(hf), ptr(up[0]), ptr(up[1]), ptr(up[2]), ptr(gwf_in), ptr(gres_in),
                                       ptr(gh_in), ptr(gp), b, solver._stream()), "hn_step_backward")
        grads, off = [], 0
        for shape in ctx.param_shapes:
            cnt = int(np.prod(shape)) if len(shape) else 1
            grads.append(gp[off: off + cnt].view(shape))
            off += cnt
        return (None, gwf_in, None, gres_in, gh_in, *grads)

# --------------------------------------------------------------------------------------------------
class IterativeSolver(nn.Module):
    def __init__(self, domain_size: int, k: float, omega: float, PMLsize: int, sigma_max: float, source_location: list,
                 train_data_path: str = None, validation_data_path: str = None, test_data_path: str = None,
                 activation_function: str = "relu", architecture: str = "custom_unet", gradient_clip_val: int = 0,
                 batch_size: int = 24, buffer_size: int = 1000, depth: int = 4, features: int = 8,
                 learning_rate: float = 1e-4, loss: str = "mse", minimum_learning_rate: float = 1e-4,
                 optimizer: str = "adam", weight_decay: float = 0.0, max_iterations: int = 100,
                 source_amplitude: int = 10, source_phase: int = 0, source_smoothing: bool = False,
                 state_channels: int = 2, state_depth: int = 4, unrolling_steps: int = 10, _backend=None):
        super().__init__()
        frame = inspect.currentframe()
        names = inspect.getargvalues(frame).args
        self.hparams = HParams({n: frame.f_locals[n] for n in names if n not in ("self", "_backend")})
        self._backend = _backend            # HelmnetLib, resolved lazily so that CPU-side construction works
        self._ctx = None
        self._ctx_key = None
        self._ctx_max_batch = 0
        self._weights_dirty = True
        self._source_dirty = True
        self._engine = int(os.environ.get("HELMNET_ENGINE", "2"))   # 2: tcgen05 + fused DoubleConvs (default), 1: tcgen05, 0: fp32 CUDA cores
        self.register_buffer("sigmas", None)
        self.set_laplacian()
        self.setup_source()
        self.init_f()

        def weights_init(m):
            if isinstance(m, nn.Conv2d):
                torch.nn.init.xavier_normal_(m.weight, gain=0.02)

        self.f.apply(weights_init)

    # ---- construction ----------------------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, hparams_file=None, strict: bool = True, **kwargs):
        ckpt = load_checkpoint(checkpoint_path, map_location="cpu")
        hp = dict(ckpt["hyper_parameters"])
        hp.update(kwargs)
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k_: v for k_, v in hp.items() if k_ in accepted})
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        if map_location is not None:
            model.to(map_location)
        return model

    _DERIVED_PREFIXES = ("Lap.", "source_module.", "metric.", "replaybuffer.")

    def load_state_dict(self, state_dict, strict: bool = True):
        own = {k_ for k_ in self.state_dict().keys() if not k_.startswith(self._DERIVED_PREFIXES)}
        own.discard("sigmas")
        picked = {k_: v for k_, v in state_dict.items() if k_ in own}
        missing = sorted(own - set(picked))
        full = set(self.state_dict().keys()) | {"sigmas"}
        unexpected = sorted(k_ for k_ in state_dict if k_ not in full)   # e.g. the stale Lap.gamma_x of the shipped ckpt
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing}, unexpected {unexpected}")
        if "source" in picked and picked["source"].shape != self.source.shape:
            self.source = nn.Parameter(picked.pop("source").clone(), requires_grad=False)
        res = super().load_state_dict(picked, strict=False)
        self._weights_dirty = True
        self._source_dirty = True
        return res

    def init_f(self):
        if self.hparams.architecture != "custom_unet":
            raise NotImplementedError("Unknown architecture {}".format(self.hparams.architecture))
        self.f = HybridNet(activation_function=self.hparams.activation_function, depth=self.hparams.depth,
                           domain_size=self.hparams.domain_size, features=self.hparams.features, inchannels=6,
                           state_channels=self.hparams.state_channels, state_depth=self.hparams.state_depth)
        self.f._owner = weakref.ref(self)

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._weights_dirty = True
        self._source_dirty = True
        return out

    # ---- domain / operator / source setup (reference hybridnet.py:92-170) ------------------------------
    def set_domain_size(self, domain_size, source_location=None, source_map=None):
        self.hparams.domain_size = domain_size
        self.f.domain_size = domain_size
        self.set_laplacian()
        self.setup_source()
        self.Lap.to(self.device)
        if self.source_module is not None:
            self.source_module.to(self.device)
        if source_location is not None:
            self.set_multiple_sources([source_location])
        else:
            self.set_source_maps(source_map)
        self.f.init_by_size()
        for enc, size in zip(self.f.enc, self.f.states_dimension):
            enc.domain_size = size

    def set_laplacian(self):
        self.Lap = SpectralLaplacian(self.hparams.domain_size, self.hparams.PMLsize, self.hparams.k, self.hparams.sigma_max,
                                     owner=self)
        sx, sy = self.Lap.sigmas()
        self.sigmas = torch.stack([sx, sy]).float().to(self.device)
        self._release_ctx()

    def setup_source(self):
        n, loc = self.hparams.domain_size, self.hparams.source_location
        if not (0 <= loc[0] < n and 0 <= loc[1] < n):
            # the reference raises IndexError here when the checkpoint's default location lies outside a
            # smaller domain; the default map is overwritten by set_domain_size anyway, so use the centre.
            loc = [n // 2, n // 2]
        self.source_module = SourceModule(image_size=n, omega=self.hparams.omega, location=loc,
                                          amplitude=self.hparams.source_amplitude, phase=self.hparams.source_phase,
                                          smooth=self.hparams.source_smoothing).to(self.device)
        with torch.no_grad():
            self.set_source()

    def set_source_maps(self, sourceval):
        if sourceval is None:
            raise ValueError("set_domain_size needs source_location or source_map")
        self.source = nn.Parameter(sourceval.to(self.device), requires_grad=False)
        self._source_dirty = True

    def set_source(self):
        self.set_source_maps(self.source_module.spatial_map(0).permute(0, 3, 1, 2))

    def reset_source(self):
        with torch.no_grad():
            if not self.source_module.get_location() == self.hparams.source_location:
                self.source_module.set_new_location(self.hparams.source_location)
                self.set_source()

    def set_multiple_sources(self, source_locations):
        """hybridnet.py:161-170: one point-source map per location.  One CUDA kernel writes all S maps (hn_point_sources: the
        delta / Blackman-smoothed map of SourceModule in closed form) instead of S passes through torch.fft; the source module
        is left at the last location, as in the reference."""
        locs = [[int(l[0]), int(l[1])] for l in source_locations]
        if not locs:
            raise ValueError("set_multiple_sources needs at least one location")
        n = self.hparams.domain_size
        for r, c in locs:
            if not (0 <= r < n and 0 <= c < n):
                raise IndexError(f"source location [{r}, {c}] is outside the {n} x {n} domain")
        dev, lib = self.device, self.lib
        if lib.requires_cuda and dev.type != "cuda":
            # host-side construction (solver still on the CPU, e.g. before .to('cuda')): plain PyTorch setup as the reference does
            maps = []
            with torch.no_grad():
                for loc in locs:
                    self.source_module.set_new_location(loc)
                    maps.append(self.source_module.spatial_map(0).permute(0, 3, 1, 2))
                self.set_source_maps(torch.cat(maps, 0))
            return
        hp = self.hparams
        loc_t = torch.tensor(locs, dtype=torch.int32, device=dev)
        out = torch.empty(len(locs), 2, n, n, device=dev)
        idx = dev.index if dev.index is not None else (torch.cuda.current_device() if dev.type == "cuda" else 0)
        lib.check(lib.hn_point_sources(idx, n, len(locs), self._ptr(loc_t), float(hp.source_amplitude), float(hp.source_phase),
                                       1 if hp.source_smoothing else 0, self._ptr(out), self._stream()), "hn_point_sources")
        with torch.no_grad():
            self.source_module.set_new_location(locs[-1])
        self._loc_keepalive = loc_t
        self.set_source_maps(out)

    # ---- CUDA context management -----------------------------------------------------------------------
    @property
    def lib(self):
        if self._backend is None:
            self._backend = _lib.default_lib()
        return self._backend

    def _release_ctx(self):
        if getattr(self, "_ctx", None) is not None:
            self.lib.hn_destroy(self._ctx)
        self._ctx, self._ctx_key, self._ctx_max_batch = None, None, 0

    def __del__(self):
        try:
            self._release_ctx()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.lib.stream_for(self.device)) if self.lib.requires_cuda else C.c_void_p(0)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _prep(self, t: torch.Tensor, name: str) -> torch.Tensor:
        t = t.detach()
        if t.device != self.device:
            t = t.to(self.device)
        self.lib.check_tensor(t, name)
        return t.float().contiguous()

    def _ensure_ctx(self, batch: int, need_weights: bool = True, need_source: bool = True):
        lib, dev = self.lib, self.device
        if lib.requires_cuda and dev.type != "cuda":
            raise _lib.HelmnetError("IterativeSolver is on the CPU: call solver.to('cuda:0'); there is no CPU path")
        hp = self.hparams
        key = (str(dev), int(hp.domain_size), int(hp.PMLsize), float(hp.sigma_max), float(hp.k), float(hp.omega))
        # the context must hold the current source maps as well ([S,2,N,N], S may exceed this call's batch: the reference
        # broadcasts a [1,...] field against an [S,...] source)
        src_b = int(self.source.shape[0]) if getattr(self, "source", None) is not None and self.source.dim() == 4 else 1
        want = max(batch, src_b, 1)
        if self._ctx is None or key != self._ctx_key or want > self._ctx_max_batch:
            self._release_ctx()
            max_batch = want
            ctx = C.c_void_p()
            idx = dev.index if dev.index is not None else (torch.cuda.current_device() if dev.type == "cuda" else 0)
            lib.check(lib.hn_create(C.byref(ctx), idx, hp.domain_size, max_batch, hp.PMLsize, float(hp.sigma_max),
                                    float(hp.k), float(hp.omega)), "hn_create")
            self._ctx, self._ctx_key, self._ctx_max_batch = ctx, key, max_batch
            self._weights_dirty = self._source_dirty = True
            if lib.requires_cuda:
                lib.check(lib.hn_set_engine(ctx, self._engine), "hn_set_engine")
        if self._weights_dirty and need_weights:
            blob = self.f.weight_blob()
            lib.check(lib.hn_load_weights(self._ctx, self._ptr(blob), blob.numel()), "hn_load_weights")
            self._weights_dirty = False
        if self._source_dirty and need_source:
            src = self.source.detach()
            if src.device != dev:
                src = src.to(dev)
            if src.dtype != torch.float32:
                src = src.float()
            n = hp.domain_size
            if src.dim() != 4 or src.shape[1] != 2 or src.shape[2] != n or src.shape[3] != n:
                raise ValueError(f"source must be [S,2,{n},{n}], got {tuple(src.shape)}")
            lib.check_tensor(src, "source")
            strides = (C.c_int64 * 4)(*src.stride())
            lib.check(lib.hn_set_source(self._ctx, self._ptr(src), src.shape[0], strides, self._stream()), "hn_set_source")
            self._source_keepalive = src
            self._source_dirty = False
        return self._ctx

    def set_engine(self, engine: int):
        """0: fp32 CUDA-core convolutions; 1: tcgen05 split-fp16 tensor-core convolutions, one kernel per conv;
        2 (default): the same with every DoubleConv of a level that is 8..256 pixels wide fused into one kernel."""
        self._engine = int(engine)
        if self._ctx is not None:
            self.lib.check(self.lib.hn_set_engine(self._ctx, self._engine), "hn_set_engine")

    def sync_check(self):
        """Synchronise and raise if a kernel recorded a device-side fault."""
        if self._ctx is not None:
            self.lib.check(self.lib.hn_sync_check(self._ctx, self._stream()), "hn_sync_check")

    def sync_weights(self):
        """Call after mutating ``solver.f`` parameters in place."""
        self._weights_dirty = True

    # ---- reference API: pieces of the loop -----------------------------------------------------------------
    @staticmethod
    def test_loss_function(x):
        return x.pow(2).mean((1, 2, 3)).sqrt()

    def get_initials(self, sos_maps: torch.Tensor):
        k_sq = (self.hparams.omega / sos_maps) ** 2
        wavefield = torch.zeros(k_sq.shape[0], 2, k_sq.shape[2], k_sq.shape[3], device=k_sq.device)
        return k_sq, wavefield

    def apply_laplacian(self, x: torch.Tensor):
        x = self._prep(x, "x")
        ctx = self._ensure_ctx(x.shape[0], need_weights=False, need_source=False)   # L(u) needs neither
        out = torch.empty_like(x)
        self.lib.check(self.lib.hn_laplacian(ctx, self._ptr(x), self._ptr(out), x.shape[0], self._stream()), "hn_laplacian")
        return out

    def get_residual(self, x: torch.Tensor, k_sq: torch.Tensor):
        x = self._prep(x, "x")
        k_sq = self._prep(k_sq, "k_sq")
        s_b = int(self.source.shape[0])
        if x.shape[0] == 1 and s_b > 1:      # hybridnet.py:556 broadcasts a single field against S source maps
            x = x.expand(s_b, -1, -1, -1).contiguous()
        if k_sq.shape[0] == 1 and x.shape[0] > 1:
            k_sq = k_sq.expand(x.shape[0], -1, -1, -1).contiguous()
        if s_b not in (1, x.shape[0]):
            raise ValueError(f"source holds {s_b} maps but the field batch is {x.shape[0]}: they must match (or one of them be 1)")
        ctx = self._ensure_ctx(x.shape[0])
        out = torch.empty_like(x)
        self.lib.check(self.lib.hn_residual(ctx, self._ptr(x), self._ptr(k_sq), self._ptr(out), self._ptr(None), x.shape[0],
                                            self._stream()), "hn_residual")
        return out

    def _push_states(self, ctx, batch):
        hs = self.f.get_states()
        if any(h is None for h in hs):
            raise ValueError("You must set or clear the state before using this module")
        flat = self._prep(self.f.flatten_state(hs), "state")
        if flat.shape[0] != batch:
            raise ValueError("hidden state batch does not match the input batch")
        return flat

    def _pull_states(self, ctx, batch):
        flat = torch.empty(batch, 2, self.f.total_state_length, device=self.device)
        self.lib.check(self.lib.hn_get(ctx, self._ptr(None), self._ptr(None), self._ptr(flat), self._stream()), "hn_get")
        self.f.set_states(flat, flatten=True)

    def _unet_forward(self, x: torch.Tensor):
        x = self._prep(x, "input")
        b = x.shape[0]
        ctx = self._ensure_ctx(b)
        flat = self._push_states(ctx, b)
        lib = self.lib
        lib.check(lib.hn_set_state(ctx, self._ptr(None), self._ptr(None), self._ptr(None), self._ptr(flat), b, self._stream()),
                  "hn_set_state")
        out = torch.empty(b, 2, x.shape[2], x.shape[3], device=self.device)
        lib.check(lib.hn_unet(ctx, self._ptr(x), self._ptr(out), b, self._stream()), "hn_unet")
        flat2 = torch.empty_like(flat)
        lib.check(lib.hn_get_states(ctx, self._ptr(flat2), b, self._stream()), "hn_get_states")
        self.f.set_states(flat2, flatten=True)
        return out

    def _run(self, ctx, batch, num_iterations, return_wavefields, return_states, return_residuals):
        n, dev, lib = self.hparams.domain_size, self.device, self.lib
        k_it = int(num_iterations)
        rmse = torch.empty(k_it, batch, device=dev)
        wf_hist = torch.empty(k_it, batch, 2, n, n, device=dev) if return_wavefields else None
        res_hist = torch.empty(k_it, batch, 2, n, n, device=dev) if return_residuals else None
        h_hist = torch.empty(k_it, batch, 2, self.f.total_state_length, device=dev) if return_states else None
        lib.check(lib.hn_run(ctx, k_it, self._ptr(rmse), self._ptr(wf_hist), self._ptr(res_hist), self._ptr(h_hist),
                             self._stream()), "hn_run")
        return rmse, wf_hist, res_hist, h_hist

    def _collect(self, ctx, batch, rmse, wf_hist, res_hist, h_hist, last_iteration):
        n, dev, lib = self.hparams.domain_size, self.device, self.lib
        wf_last = None if wf_hist is not None else torch.empty(batch, 2, n, n, device=dev)
        res_last = None if res_hist is not None else torch.empty(batch, 2, n, n, device=dev)
        flat = torch.empty(batch, 2, self.f.total_state_length, device=dev)
        lib.check(lib.hn_get(ctx, self._ptr(wf_last), self._ptr(res_last), self._ptr(flat), self._stream()), "hn_get")
        self.f.set_states(flat, flatten=True)
        return {
            "wavefields": list(wf_hist.unbind(0)) if wf_hist is not None else [wf_last],
            "residuals": list(res_hist.unbind(0)) if res_hist is not None else [res_last],
            "states": list(h_hist.unbind(0)) if h_hist is not None else [],
            "last_iteration": last_iteration,
            "residual_rmse": rmse,
        }

    # ---- test-set evaluation hooks (reference hybridnet.py:299-330, driven by evaluate.py:27-29) ------------------
    def test_step(self, batch, batch_idx=0):
        self.reset_source()
        output = self.forward(batch, num_iterations=self.hparams.max_iterations, return_wavefields=True, return_states=False,
                              return_residuals=False)
        return {"losses": output["residual_rmse"].transpose(0, 1).contiguous(), "wavefields": list(output["wavefields"])}

    def test_epoch_end(self, outputs, out_dir: str = "results"):
        """Writes the two files produce_figures.py:51-64 consumes."""
        import os
        os.makedirs(out_dir, exist_ok=True)
        all_losses = torch.cat([o["losses"] for o in outputs], dim=0).cpu().numpy()
        np.save(os.path.join(out_dir, "evolution_of_model_RMSE_on_test_set"), all_losses)
        wavefields = torch.cat([torch.stack(o["wavefields"], 0) for o in outputs], 1).permute(1, 0, 2, 3, 4)
        np.save(os.path.join(out_dir, "evolution_of_wavefields_on_test_set"), wavefields.cpu().numpy())
        return all_losses

    def single_step(self, wavefield, k_sq, residual, get_residual: bool = True):
        out = self.n_steps(wavefield, k_sq, residual, 1)
        if get_residual:
            return out["wavefields"][0], out["residuals"][0]
        return out["wavefields"][0]

    def _wants_grad(self, *tensors):
        if not torch.is_grad_enabled():
            return False
        if any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
            return True
        return any(p.requires_grad for p in self.f.parameters())

    def _n_steps_autograd(self, wavefield, k_sq, residual, num_iterations, return_wavefields, return_states):
        """The training unroll (reference hybridnet.py:586-623 with autograd recording): every step is one _SingleStepFn node,
        so ``loss.backward()`` on the returned residuals / wavefields / states reaches the inputs and ``solver.f``'s parameters."""
        params = list(self.f.parameters())
        versions = tuple(p._version for p in params)
        if versions != getattr(self, "_param_versions", None):     # an optimizer step mutated the parameters in place
            self._param_versions = versions
            self._weights_dirty = True
        hs = self.f.get_states()
        if any(h is None for h in hs):
            raise ValueError("You must set or clear the state before using this module")
        hflat = self.f.flatten_state(hs)
        wavefields, residuals, states = [], [], []
        for _ in range(num_iterations):
            wavefield, residual, hflat = _SingleStepFn.apply(self, wavefield, k_sq, residual, hflat, *params)
            self.f.set_states(hflat, flatten=True)
            residuals.append(residual)
            if return_wavefields:
                wavefields.append(wavefield)
            if return_states:
                states.append(self.f.get_states(flatten=True))
        if not return_wavefields:
            wavefields.append(wavefield)
        with torch.no_grad():
            rmse = torch.stack([self.test_loss_function(r) for r in residuals])
        return {"wavefields": wavefields, "residuals": residuals, "states": states, "last_iteration": num_iterations - 1,
                "residual_rmse": rmse}

    def n_steps(self, wavefield, k_sq, residual, num_iterations, return_wavefields=False, return_states=False,
                return_residuals=True):
        if num_iterations < 1:
            raise ValueError("num_iterations must be >= 1")
        if self._wants_grad(wavefield, residual, *[h for h in self.f.get_states() if h is not None]):
            return self._n_steps_autograd(wavefield, k_sq, residual, num_iterations, return_wavefields, return_states)
        wavefield, k_sq, residual = self._prep(wavefield, "wavefield"), self._prep(k_sq, "k_sq"), self._prep(residual, "residual")
        b = wavefield.shape[0]
        ctx = self._ensure_ctx(b)
        flat = self._push_states(ctx, b)
        self.lib.check(self.lib.hn_set_state(ctx, self._ptr(wavefield), self._ptr(residual), self._ptr(k_sq), self._ptr(flat), b,
                                             self._stream()), "hn_set_state")
        hist = self._run(ctx, b, num_iterations, return_wavefields, return_states, return_residuals)
        return self._collect(ctx, b, *hist, last_iteration=num_iterations - 1)

    def forward(self, sos_maps, return_wavefields=False, return_states=False, num_iterations=None, stop_if_diverge=False,
                return_residuals=True):
        """Reference hybridnet.py:654-697.  ``residuals`` is the per-iteration list of residual tensors as in
        the reference; pass ``return_residuals=False`` for large runs to keep only the last one --
        ``residual_rmse`` ([iterations, batch], the test_loss_function of every residual) is always returned."""
        if num_iterations is None:
            num_iterations = self.hparams.max_iterations
        if num_iterations < 1:
            raise ValueError("num_iterations must be >= 1")
        sos = self._prep(sos_maps, "sos_maps")
        n = self.hparams.domain_size
        if sos.dim() != 4 or sos.shape[1] != 1 or sos.shape[2] != n or sos.shape[3] != n:
            raise ValueError(f"sos_maps must be [B,1,{n},{n}], got {tuple(sos.shape)}")
        b = sos.shape[0]
        ctx = self._ensure_ctx(b)
        self.lib.check(self.lib.hn_reset(ctx, self._ptr(sos), b, self._stream()), "hn_reset")
        hist = self._run(ctx, b, num_iterations, return_wavefields, return_states, return_residuals)
        return self._collect(ctx, b, *hist, last_iteration=num_iterations - 1)

    def forward_variable_src(self, sos_maps, src_time_pairs, return_wavefields=False, return_states=False,
                             num_iterations=None, stop_if_diverge=False, return_residuals=True):
        """Reference hybridnet.py:699-754: swap the source map at given iterations and recompute the residual."""
        if num_iterations is None:
            num_iterations = self.hparams.max_iterations
        times = sorted(set(int(t) for t in src_time_pairs["iteration"] if 0 <= int(t) < num_iterations))
        src_maps = iter(src_time_pairs["src_maps"])
        sos = self._prep(sos_maps, "sos_maps")
        b, n, dev, lib = sos.shape[0], self.hparams.domain_size, self.device, self.lib
        ctx = self._ensure_ctx(b)
        lib.check(lib.hn_reset(ctx, self._ptr(sos), b, self._stream()), "hn_reset")
        cuts = sorted(set([0] + times + [num_iterations]))
        parts = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            if lo in times:
                self.set_source_maps(next(src_maps))
                ctx = self._ensure_ctx(b)
                wf = torch.empty(b, 2, n, n, device=dev)
                lib.check(lib.hn_get(ctx, self._ptr(wf), self._ptr(None), self._ptr(None), self._stream()), "hn_get")
                res = torch.empty_like(wf)
                lib.check(lib.hn_residual(ctx, self._ptr(wf), self._ptr(None), self._ptr(res), self._ptr(None), b, self._stream()),
                          "hn_residual")
                lib.check(lib.hn_set_state(ctx, self._ptr(None), self._ptr(res), self._ptr(None), self._ptr(None), b, self._stream()),
                          "hn_set_state")
            if hi > lo:
                parts.append(self._run(ctx, b, hi - lo, return_wavefields, return_states, return_residuals))
        cat = [torch.cat([p[i] for p in parts], 0) if parts[0][i] is not None else None for i in range(4)]
        return self._collect(ctx, b, *cat, last_iteration=num_iterations - 1)
