"""Seeded synthetic sound-speed maps for tests and benchmarks.

A numpy-only generator with the shape statistics of the reference's training data
(helmnet/dataloaders.py:83-156: a random thick elliptical "skull" shell, sos in [1, 2]) plus a smooth
low-pass heterogeneity inside the shell.  Input generator only -- not part of the parity surface.
"""
import numpy as np
import torch


def synthetic_sos(batch: int, n: int, seed: int = 0, contrast: float = 1.0) -> torch.Tensor:
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    out = np.ones((batch, 1, n, n), np.float32)
    for b in range(batch):
        cx, cy = n / 2 + rng.randn(2) * n * 0.03
        a = n * (0.22 + 0.1 * rng.rand())
        bb = n * (0.22 + 0.1 * rng.rand())
        th = rng.rand() * np.pi
        thick = max(2.0, n * (0.02 + 0.03 * rng.rand()))
        xr = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        yr = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        rho = np.sqrt((xr / a) ** 2 + (yr / bb) ** 2)
        shell = np.abs(rho - 1.0) * min(a, bb) < thick / 2
        boost = contrast * (0.5 + 0.5 * rng.rand())
        m = np.ones((n, n))
        m[shell] += boost
        # smooth heterogeneity: low-pass filtered noise, +-0.1
        noise = rng.randn(n, n)
        kx = np.fft.fftfreq(n)[None, :]
        ky = np.fft.fftfreq(n)[:, None]
        filt = np.exp(-0.5 * (kx ** 2 + ky ** 2) * (2 * np.pi * 8.0) ** 2)
        sm = np.real(np.fft.ifft2(np.fft.fft2(noise) * filt))
        sm = 0.1 * sm / (np.abs(sm).max() + 1e-12)
        inside = rho < 1.0
        m[inside & ~shell] += sm[inside & ~shell]
        out[b, 0] = np.clip(m, 1.0, 2.0)
    return torch.from_numpy(out)
