"""TEST INFRASTRUCTURE — minimal stand-in for `lightning` (absent from this image).

Only what the reference's hot-path modules touch at import / call time
(`models/base/model.py:15-25,40-42`, `core/training/datamodule.py:36`,
`core/training/trainer.py:3-5`): `LightningModule` is an `nn.Module` with no-op
`save_hyperparameters` / `log`; the Trainer/DataModule names exist so that
`core/training/__init__.py` imports.  Nothing here is on the product path.
"""
import torch.nn as _nn


class LightningModule(_nn.Module):
    def save_hyperparameters(self, *args, **kwargs):
        return None

    def log(self, *args, **kwargs):
        return None

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            import torch

            return torch.device("cpu")


class LightningDataModule:
    def __init__(self, *args, **kwargs):
        pass


class Trainer:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("lightning shim: Trainer is not available (out of scope, SURVEY.md #4)")
