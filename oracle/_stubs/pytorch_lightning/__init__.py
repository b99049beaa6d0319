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

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

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
