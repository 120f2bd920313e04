"""`LightningModule` for NeRFNetwork: the real pytorch_lightning class when it is installed (the reference
pins 1.5.10), otherwise a small nn.Module stand-in with the handful of members the reference touches
(`log`, `device`, `load_from_checkpoint`, `logger`, `trainer`).  Checkpoints keep PL's dict layout either way."""
import torch
import torch.nn as nn

try:                                        # pragma: no cover - not installable offline
    from pytorch_lightning import LightningModule  # noqa: F401
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()
            self.logged = {}
            self.trainer = None
            self.logger = None

        def log(self, name, value, **kwargs):
            # detached: a logged loss must not keep the step's autograd graph (and the parameters' grad accumulators, which
            # remember the stream they were created on) alive into the next step - a CUDA-graph capture of that step would inherit
            # accumulators bound to the warm-up stream
            value = value.detach() if isinstance(value, torch.Tensor) else value
            self.logged[name] = value
            if self.trainer is not None and hasattr(self.trainer, "record"):
                self.trainer.record(name, value)

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
            ckpt = torch.load(str(checkpoint_path), map_location=map_location or "cpu", weights_only=False)
            hparams = dict(ckpt.get("hyper_parameters", {}))
            hparams.update(kwargs)
            model = cls(**hparams)
            model.load_state_dict(ckpt["state_dict"])
            return model
