import torch
from torch import nn


class MeanAbsoluteError(nn.Module):
    def forward(self, a, b):
        return (a - b).abs().mean()
