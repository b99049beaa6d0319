"""Monochromatic point source map -- setup-time only, plain PyTorch on the solver's device.

Mirrors the interface of the reference's ``SourceModule`` (helmnet/source_module.py:4-116): a delta of
``amplitude`` at ``location`` (row, col), optionally smoothed with a Blackman window in the spatial
frequency domain, returned as real/imag channels.  The map is an *input* of the CUDA hot path
(SURVEY.md section 8, row f1), not part of it.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class SourceModule(nn.Module):
    def __init__(self, image_size: int, omega: float = 1, location=(180, 50), amplitude: float = 1.0,
                 phase: float = 0.0, smooth: bool = True):
        super().__init__()
        self.L = image_size
        self.location = list(location)
        self.omega = omega
        self.amplitude = amplitude
        self.phase = phase
        self.smooth = smooth
        self.register_buffer("_dummy_for_device", torch.tensor(1))
        self.register_buffer("_abs_spatial_map", None)
        self.make_abs_spatial_map(smooth=smooth)

    def make_abs_spatial_map(self, smooth: bool = True):
        dev = self._dummy_for_device.device
        delta = torch.zeros((self.L, self.L), device=dev)
        delta[self.location[0], self.location[1]] = self.amplitude
        spec = torch.fft.fftshift(torch.fft.fft2(delta))
        if smooth:
            win = torch.blackman_window(self.L, device=dev)
            spec = spec * torch.outer(win, win)
        self._abs_spatial_map = torch.abs(torch.fft.ifft2(torch.fft.ifftshift(spec)))

    def set_new_location(self, location):
        if self.location[0] != location[0] or self.location[1] != location[1]:
            self.location = list(location)
            self.make_abs_spatial_map(smooth=self.smooth)

    def get_location(self):
        return self.location

    def spatial_map(self, t: float) -> torch.Tensor:
        """[1, L, L, 2] real/imag map at time t."""
        arg = torch.tensor(self.omega * t + self.phase, device=self._dummy_for_device.device)
        with torch.no_grad():
            out = torch.stack([self._abs_spatial_map * torch.cos(arg), self._abs_spatial_map * torch.sin(arg)], dim=2)
        return out.unsqueeze(0)
