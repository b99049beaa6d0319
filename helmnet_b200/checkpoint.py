"""Checkpoint loading without PyTorch-Lightning.

The shipped ``trained_models/jcp_paper_trained_weights.ckpt`` is a legacy (non-zip) pickle written by
pytorch-lightning 0.9.0; unpickling it needs ``pytorch_lightning.utilities.parsing.AttributeDict``
(SURVEY.md Appendix B).  Lightning is not a dependency here, so a stand-in module is installed in
``sys.modules`` for the duration of ``torch.load`` when the real one is not importable.

Replaces the inherited ``LightningModule.load_from_checkpoint`` used at the reference's
README.md:58-60 / evaluate.py:51.
"""
from __future__ import annotations

import contextlib
import importlib.util
import sys
import types
from typing import Any, Dict

import torch


class HParams(dict):
    """Attribute-style dict, the role Lightning's AttributeDict plays for ``solver.hparams``."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]


@contextlib.contextmanager
def _lightning_standin():
    if importlib.util.find_spec("pytorch_lightning") is not None:
        yield
        return
    names = ["pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.parsing"]
    mods = {n: types.ModuleType(n) for n in names}
    mods[names[0]].utilities = mods[names[1]]
    mods[names[1]].parsing = mods[names[2]]
    mods[names[2]].AttributeDict = HParams
    saved = {n: sys.modules.get(n) for n in names}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for n in names:
            if saved[n] is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = saved[n]


def load_checkpoint(path: str, map_location: Any = "cpu") -> Dict[str, Any]:
    """Returns the raw checkpoint dict (``hyper_parameters``, ``state_dict``, ...)."""
    with _lightning_standin():
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if not isinstance(ckpt, dict) or "state_dict" not in ckpt:
        raise ValueError(f"{path}: not a helmnet checkpoint (no 'state_dict')")
    ckpt["hyper_parameters"] = HParams(dict(ckpt.get("hyper_parameters", {})))
    return ckpt
