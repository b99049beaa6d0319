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


# --------------------------------------------------------------------------------------------------
# the synthetic inputs of the BASELINE.json configurations.
#
# `skull_outline_map` restates the recipe of the reference's training-set generator
# (helmnet/dataloaders.py:83-156, EllipsesDataset._make_ellipsoid): a closed curve made of four harmonics with random
# amplitudes / phases, drawn as a polyline of random thickness; sos = background + outline * (boost_min + U * boost_rand).
# The random numbers are drawn from `rng` in the recipe's order (4 + 4 amplitudes, 4 + 4 phases, thickness, boost), so
# with `rng = np.random` after np.random.seed(s) it produces the very map the reference function produces
# (tests/test_checkpoint_and_api.py checks that whenever /root/reference is present).  Input generator only -- not part
# of the parity surface.
# --------------------------------------------------------------------------------------------------
_AMP_MEAN = np.array([1.0, 0.0, 0.0, 0.0])
_AMP_STD = np.array([0.1, 0.05, 0.025, 0.01])
CONFIG_SPECS = {"C2": (96, 0), "C3": (256, 1), "C4": (512, 2), "C5": (1024, 3)}   # name -> (domain size, seed of 8(d))


def _draw_polyline(n: int, xs: np.ndarray, ys: np.ndarray, thickness: int) -> np.ndarray:
    """Closed polyline through the (truncated-to-integer) vertices with the given stroke thickness -> {0,1} mask [n, n]."""
    try:
        import cv2
    except ImportError as e:  # pragma: no cover - the image ships opencv
        raise RuntimeError("the 8(d) synthetic maps rasterise their outline with cv2.polylines (opencv is part of the image)") from e
    canvas = np.zeros((n, n, 3), dtype=np.uint8)
    verts = np.array([xs, ys], np.int32).T[None]
    cv2.polylines(canvas, [verts], True, (1, 0, 0), thickness=thickness)
    return canvas[:, :, 0].astype(np.float32)


def skull_outline_map(n: int, rng, avg_thickness: float = 2, std_thickness: float = 8, background: float = 1.0,
                      boost_min: float = 0.5, boost_rand: float = 0.5, phase_std: float = np.pi / 16) -> np.ndarray:
    theta = np.linspace(0.0, 2.0 * np.pi, num=360, endpoint=True)
    amp_x = _AMP_MEAN + rng.randn(4) * _AMP_STD
    amp_y = _AMP_MEAN + rng.randn(4) * _AMP_STD
    ph_x = rng.randn(4) * phase_std
    ph_y = rng.randn(4) * phase_std
    cx = np.zeros_like(theta)
    cy = np.zeros_like(theta)
    for h in range(4):                     # same summation order as the recipe (bit-identical vertices)
        cx = cx + np.sin(theta * (h + 1) + ph_x[h]) * amp_x[h]
        cy = cy + np.cos(theta * (h + 1) + ph_y[h]) * amp_y[h]
    cx, cy = (cx + 2) / 4 * n, (cy + 2) / 4 * n
    thickness = int(avg_thickness + rng.rand(1)[0] * std_thickness)
    mask = _draw_polyline(n, cx, cy, thickness)
    boost = rng.rand(1) * boost_rand + boost_min
    return background + mask * boost       # float64 [n, n], as the recipe returns it


def _smooth_heterogeneity(n: int, rng, sigma_px: float = 8.0, amplitude: float = 0.1) -> np.ndarray:
    from scipy.ndimage import gaussian_filter
    g = gaussian_filter(rng.randn(n, n), sigma=sigma_px, mode="wrap")
    return amplitude * g / (np.abs(g).max() + 1e-12)


def _config_map(config: str, rng) -> np.ndarray:
    if config == "C2":
        return skull_outline_map(96, rng)
    if config == "C3":
        return skull_outline_map(256, rng, avg_thickness=5, std_thickness=21) + _smooth_heterogeneity(256, rng)
    if config == "C4":
        return skull_outline_map(512, rng, avg_thickness=10, std_thickness=40, boost_min=0.9, boost_rand=0.1)
    if config == "C5":
        m = np.empty((1024, 1024))
        for ty in range(4):
            for tx in range(4):
                m[ty * 256:(ty + 1) * 256, tx * 256:(tx + 1) * 256] = _config_map("C3", rng)
        return m
    raise ValueError(f"unknown configuration {config!r} (C2, C3, C4 or C5)")


def config_sos(config: str, count: int, seed=None, start: int = 0) -> torch.Tensor:
    """Sound-speed maps of BASELINE.json's configurations as SURVEY.md 8(d) defines them: [count, 1, N, N] float32 in [1, 2].

      "C2"  96^2   : `_make_ellipsoid(imsize=96)` (training-domain maps), seed 0
      "C3"  256^2  : `_make_ellipsoid(256, avg_thickness=5, std_thickness=21)` + Gaussian-filtered N(0,1) (sigma 8 px) scaled
                     to +-0.1, clipped to [1, 2], seed 1                      <- the bench.py workload (256 distinct maps)
      "C4"  512^2  : `_make_ellipsoid(512, 10, 40, minimal_skull_sos_boost=0.9, maximal_random_skull_boost=0.1)`, seed 2
      "C5"  1024^2 : 4 x 4 tiling of 256^2 C3-style maps, seed 3
    Returns maps start .. start + count - 1.  Map i depends only on (config, seed, i) -- the generator is re-seeded per map --
    so every rank of a sharded run builds exactly its own slice of the batch.
    """
    if config not in CONFIG_SPECS:
        raise ValueError(f"unknown configuration {config!r} (C2, C3, C4 or C5)")
    n, default_seed = CONFIG_SPECS[config]
    seed = default_seed if seed is None else int(seed)
    out = np.empty((count, 1, n, n), np.float32)
    for i in range(count):
        rng = np.random.RandomState((seed * 100003 + start + i) % (2 ** 31 - 1))
        out[i, 0] = np.clip(_config_map(config, rng), 1.0, 2.0)
    return torch.from_numpy(out)
