"""Minimal stand-in for pytorch_lightning==1.5.10 (not installable here: no network).

Only used by tests/golden/make_golden.py to import the UNMODIFIED reference modules
from /root/reference in this container.  Not part of the product path.
"""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.logged = {}

    def log(self, name, value, **kw):
        self.logged[name] = value

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @classmethod
    def load_from_checkpoint(cls, path, **kw):
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        model = cls(**ckpt.get("hyper_parameters", {}))
        model.load_state_dict(ckpt["state_dict"])
        return model


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Trainer:
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")
