"""CPU oracle for the helmnet inference inner loop -- TEST INFRASTRUCTURE ONLY.

A functional (module-free, Lightning-free) restatement of the reference's hot
path in plain PyTorch CPU ops, used as the checker for the CUDA path.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this file; the product package
``helmnet_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against fixtures in ``tests/golden/`` that were produced by running the
unmodified reference (``/root/reference/helmnet``) in the build container via
``oracle/make_golden.py`` (the reference has no golden vectors of its own,
SURVEY.md section 4).

Third-party arithmetic under the path (torch.fft on pocketfft, conv2d on
oneDNN) is the same installed torch the reference itself calls.

Each function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# operator constants: helmnet/spectral.py:122-146 (FourierDerivative.__init__)
# and :267-363 (FastLaplacianWithPML.init_variables / get_gamma_functions)
# --------------------------------------------------------------------------- #
def wavenumbers(n: int) -> np.ndarray:
    """spectral.py:126-127 -- k vector in FFT order, float64."""
    k = 2 * np.pi * np.linspace(-0.5, 0.5, n, endpoint=False)
    return np.concatenate((k[n // 2:], k[: n // 2]))


def pml_profiles(n: int, pml: int, sigma_max: float, k0: float):
    """spectral.py:306-338 -- 1-D sigma, a(x) and b(x) (complex128).

    The reference builds 2-D meshgrids; every quantity depends on one axis
    only, so the 1-D vectors carry all the information.
    """
    coord = np.arange(pml)
    sigma_outer = sigma_max * (np.abs(1 - coord / pml) ** 2)
    sigma = np.zeros((n,))
    sigma[:pml] = sigma_outer
    sigma[-pml:] = np.flip(sigma_outer)
    inv_gamma = 1.0 / (np.ones_like(sigma) + (1j / k0) * sigma)
    sigma_prime = -2 * sigma_max * (1 - coord / pml) / pml
    sp = np.zeros((n,))
    sp[:pml] = sigma_prime
    sp[-pml:] = -np.flip(sigma_prime)
    gamma_prime = (1j / k0) * sp
    a = -gamma_prime * (inv_gamma ** 3)
    b = inv_gamma ** 2
    return sigma, a, b


def make_operator(n: int, pml: int, sigma_max: float, k0: float, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """All tensors FastLaplacianWithPML registers (spectral.py:267-363).

    Shapes follow the reference: [1, N, N, 2] real/imag pairs; x varies along
    the last spatial axis (W), y along H (np.meshgrid default 'xy' indexing,
    spectral.py:130-139).  k is cast to float32 BEFORE squaring
    (spectral.py:141, 281) -- kept here even for dtype=float64 so the fp64
    arbiter sees the same operator.
    """
    k = wavenumbers(n)
    k32 = torch.from_numpy(k).float()                      # :141 .float()
    kx2d = k32[None, :].expand(n, n)                        # varies along W
    ky2d = k32[:, None].expand(n, n)                        # varies along H
    kx = kx2d[None, :, :, None]
    ky = ky2d[None, :, :, None]
    kx_sq = kx.pow(2)                                       # :281 float32 square
    ky_sq = ky.pow(2)
    z = torch.zeros_like(kx)
    op = {
        "kx": torch.cat([z, kx], -1),                       # :284 imaginary
        "ky": torch.cat([z, ky], -1),
        "kx_sq": torch.cat([-kx_sq, z], -1),                # :286 negated
        "ky_sq": torch.cat([-ky_sq, z], -1),
    }
    sigma, a, b = pml_profiles(n, pml, sigma_max, k0)
    sx, sy = np.meshgrid(sigma, sigma)
    ax2, ay2 = np.meshgrid(a, a)
    bx2, by2 = np.meshgrid(b, b)

    def pair(c):
        return torch.stack([torch.from_numpy(np.real(c)), torch.from_numpy(np.imag(c))], -1)[None].float()

    op.update(ax=pair(ax2), bx=pair(bx2), ay=pair(ay2), by=pair(by2))
    op["sigma_x"] = torch.tensor(sx).float()
    op["sigma_y"] = torch.tensor(sy).float()
    op = {k_: v.to(dtype).contiguous() for k_, v in op.items()}
    # hybridnet.py:126-131 -- network input channels 4,5
    op["sigmas"] = torch.stack([op["sigma_x"], op["sigma_y"]], 0)
    return op


def complex_mul(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """spectral.py:6-18."""
    real = x[..., 0] * y[..., 0] - x[..., 1] * y[..., 1]
    imag = x[..., 1] * y[..., 0] + x[..., 0] * y[..., 1]
    return torch.stack([real, imag], dim=-1)


def laplacian(u: torch.Tensor, op: Dict[str, torch.Tensor]) -> torch.Tensor:
    """spectral.py:31-79 -- u is [B, H, W, 2]."""
    u_fft = torch.view_as_real(torch.fft.fftn(torch.view_as_complex(u), dim=(-2, -1), norm="backward"))
    dx = complex_mul(u_fft, op["kx"])
    dy = complex_mul(u_fft, op["ky"])
    ddx = complex_mul(u_fft, op["kx_sq"])
    ddy = complex_mul(u_fft, op["ky_sq"])
    der = torch.view_as_real(
        torch.fft.ifftn(torch.view_as_complex(torch.stack([dx, dy, ddx, ddy], 0)), dim=(-2, -1), norm="backward")
    )
    return (
        complex_mul(op["ax"], der[0])
        + complex_mul(op["ay"], der[1])
        + complex_mul(op["bx"], der[2])
        + complex_mul(op["by"], der[3])
    )


def apply_laplacian(x: torch.Tensor, op) -> torch.Tensor:
    """hybridnet.py:540-542 -- NCHW in, NCHW (permuted view) out."""
    return laplacian(x.permute(0, 2, 3, 1).contiguous(), op).permute(0, 3, 1, 2)


def get_residual(x: torch.Tensor, k_sq: torch.Tensor, source: torch.Tensor, op) -> torch.Tensor:
    """hybridnet.py:544-556."""
    return apply_laplacian(x, op) + k_sq * x - source


def rmse(x: torch.Tensor) -> torch.Tensor:
    """hybridnet.py:295-297 (test_loss_function)."""
    return x.pow(2).mean((1, 2, 3)).sqrt()


# --------------------------------------------------------------------------- #
# source term: helmnet/source_module.py:41-116, hybridnet.py:151-170
# --------------------------------------------------------------------------- #
def point_source(n: int, location: Sequence[int], amplitude: float = 10.0, phase: float = 0.0,
                 omega: float = 1.0, smooth: bool = False) -> torch.Tensor:
    """Returns the [1, 2, N, N] permuted (non-contiguous) view the reference stores."""
    spatial_map = torch.zeros((n, n))
    spatial_map[location[0], location[1]] = amplitude
    f = torch.fft.fftshift(torch.fft.fft2(spatial_map))
    if smooth:
        bw = torch.blackman_window(n)
        f = f * torch.outer(bw, bw)
    amp = torch.abs(torch.fft.ifft2(torch.fft.ifftshift(f)))
    t = torch.tensor(omega * 0 + phase)
    src = torch.stack([amp * torch.cos(t), amp * torch.sin(t)], dim=2)[None]
    return src.permute(0, 3, 1, 2)


def point_sources(n: int, locations, **kw) -> torch.Tensor:
    """hybridnet.py:161-170 (set_multiple_sources)."""
    return torch.cat([point_source(n, loc, **kw) for loc in locations], 0)


# --------------------------------------------------------------------------- #
# learned optimizer: helmnet/architectures.py
# --------------------------------------------------------------------------- #
def prelu(x: torch.Tensor, slope: torch.Tensor) -> torch.Tensor:
    """architectures.py:32-33 -- nn.PReLU() with ONE (signed) slope."""
    return F.prelu(x, slope)


def double_conv(x, w, prefix: str):
    """architectures.py:63-84 -- conv3x3 -> PReLU -> conv3x3 (no 2nd activation)."""
    x = F.conv2d(x, w[prefix + ".double_conv.0.weight"], w[prefix + ".double_conv.0.bias"], padding=1)
    x = prelu(x, w[prefix + ".double_conv.1.weight"])
    return F.conv2d(x, w[prefix + ".double_conv.2.weight"], w[prefix + ".double_conv.2.bias"], padding=1)


def unet_forward(w: Dict[str, torch.Tensor], x: torch.Tensor, states: List[torch.Tensor], depth: int = 4):
    """architectures.py:439-465 (HybridNet.forward) + :240-252 (EncoderBlock.forward).

    ``w`` holds the ``f.*`` tensors with the ``f.`` prefix stripped.
    Returns (out[B,2,N,N], new_states).
    """
    x = double_conv(x, w, "inc")
    inner = []
    new_states = []
    for d in range(depth):
        xs = torch.cat([x, states[d]], 1)
        out = double_conv(xs, w, f"enc.{d}.conv_signal")
        new_states.append(double_conv(torch.cat([out, states[d]], 1), w, f"enc.{d}.conv_state"))
        inner.append(out)
        x = F.conv2d(out, w[f"enc.{d}.down.weight"], w[f"enc.{d}.down.bias"], stride=2, padding=3)
    x = double_conv(x, w, f"decode.{depth}")
    for d in range(depth - 1, -1, -1):
        x = F.conv_transpose2d(x, w[f"up.{d}.weight"], w[f"up.{d}.bias"], stride=2, padding=3)
        x = torch.cat([x, inner[d]], 1)
        x = double_conv(x, w, f"decode.{d}")
    out = F.conv2d(x, w["outc.conv.weight"], w["outc.conv.bias"])
    return out, new_states


def zero_states(batch: int, n: int, depth: int = 4, dtype=torch.float32) -> List[torch.Tensor]:
    """architectures.py:235-238, 415-417."""
    return [torch.zeros(batch, 2, n // 2 ** d, n // 2 ** d, dtype=dtype) for d in range(depth)]


def flatten_states(states: List[torch.Tensor]) -> torch.Tensor:
    """architectures.py:419-423."""
    return torch.cat([s.reshape(s.shape[0], s.shape[1], -1) for s in states], 2)


def unflatten_states(flat: torch.Tensor, n: int, depth: int = 4) -> List[torch.Tensor]:
    """architectures.py:425-432."""
    out, off = [], 0
    for d in range(depth):
        r = n // 2 ** d
        out.append(flat[:, :, off: off + r * r].reshape(flat.shape[0], flat.shape[1], r, r))
        off += r * r
    return out


# --------------------------------------------------------------------------- #
# solver loop: helmnet/hybridnet.py:522-584, 654-697
# --------------------------------------------------------------------------- #
class Oracle:
    """Holds weights + operator tables for one domain size; all methods are @no_grad."""

    def __init__(self, weights: Dict[str, torch.Tensor], n: int, *, pml: int = 8, sigma_max: float = 2.0,
                 k0: float = 1.0, omega: float = 1.0, depth: int = 4, dtype=torch.float32):
        self.n, self.depth, self.omega, self.dtype = n, depth, omega, dtype
        self.w = {k_: v.detach().to(dtype) for k_, v in weights.items()}
        self.op = make_operator(n, pml, sigma_max, k0, dtype)
        self.source: Optional[torch.Tensor] = None

    def set_source(self, source: torch.Tensor):
        self.source = source.to(self.dtype)

    @torch.no_grad()
    def get_initials(self, sos):
        """hybridnet.py:522-538."""
        k_sq = (self.omega / sos.to(self.dtype)) ** 2
        wf = torch.zeros(k_sq.shape[0], 2, k_sq.shape[2], k_sq.shape[3], dtype=self.dtype)
        return k_sq, wf

    @torch.no_grad()
    def residual(self, wf, k_sq):
        return get_residual(wf, k_sq, self.source, self.op)

    @torch.no_grad()
    def single_step(self, wf, k_sq, res, states):
        """hybridnet.py:558-584. Returns (up_wf, new_res, new_states)."""
        sig = self.op["sigmas"].unsqueeze(0).repeat(wf.shape[0], 1, 1, 1)
        inp = torch.cat([wf, 1e3 * res, sig], dim=1)
        d, new_states = unet_forward(self.w, inp, states, self.depth)
        up = d / 1e3 + wf
        return up, self.residual(up, k_sq), new_states

    @torch.no_grad()
    def forward(self, sos, num_iterations: int, keep_wavefields: bool = False, keep_states: bool = False):
        """hybridnet.py:654-697. Returns dict with per-iteration RMSE [K,B]."""
        k_sq, wf = self.get_initials(sos)
        states = zero_states(wf.shape[0], self.n, self.depth, self.dtype)
        res = self.residual(wf, k_sq)
        hist, wfs, sts = [], [], []
        for _ in range(num_iterations):
            wf, res, states = self.single_step(wf, k_sq, res, states)
            hist.append(rmse(res))
            if keep_wavefields:
                wfs.append(wf)
            if keep_states:
                sts.append(flatten_states(states))
        return {"wavefield": wf, "residual": res, "states": states, "rmse": torch.stack(hist, 0) if hist else None,
                "wavefields": wfs, "states_hist": sts}
