"""Minimal stand-in for pytorch_lightning, TEST INFRASTRUCTURE ONLY.

Lets the unmodified reference package (/root/reference/helmnet) import and run
on CPU inside the build container so that oracle/make_golden.py can generate
golden vectors from the reference itself. Nothing in the product path imports
this. Only the pieces the reference touches are provided:
LightningModule (hparams, save_hyperparameters, device, freeze,
load_from_checkpoint) and utilities.parsing.AttributeDict (needed to unpickle
the legacy checkpoint).
"""
import inspect

import torch
from torch import nn

from .utilities.parsing import AttributeDict


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.hparams = AttributeDict()

    def save_hyperparameters(self):
        frame = inspect.currentframe().f_back
        args = inspect.getargvalues(frame)
        for name in args.args:
            if name != "self":
                self.hparams[name] = args.locals[name]

    # Lightning's DeviceDtypeModuleMixin remembers the device the module was moved to (it does not look at the parameters):
    # the reference relies on that in set_domain_size, where the freshly built source parameter still sits on the CPU when
    # `self.Lap.to(self.device)` runs (hybridnet.py:92-101).
    @property
    def device(self):
        return getattr(self, "_stub_device", torch.device("cpu"))

    def to(self, *args, **kwargs):
        out = torch._C._nn._parse_to(*args, **kwargs)
        if out[0] is not None:
            self._stub_device = out[0]
        return super().to(*args, **kwargs)

    def cuda(self, device=None):
        self._stub_device = torch.device("cuda", torch.cuda.current_device() if device is None else (device if isinstance(device, int) else torch.device(device).index or 0))
        return super().cuda(device)

    def cpu(self):
        self._stub_device = torch.device("cpu")
        return super().cpu()

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, strict=True, map_location="cpu", **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    def log(self, *a, **k):
        pass


class Trainer:  # never used by the oracle
    pass
