"""Minimal stand-in for the slice of pytorch_lightning.Trainer that the reference's train_nerf.py uses
(train_nerf.py:20-34): max_steps, per-epoch LR scheduler step, resume_from_checkpoint, validation every n epochs,
dataloader reload (the cropping switch), gradient-norm tracking, and PL-1.5.10-shaped checkpoints
(`default_root_dir/NeRF/<run>/checkpoints/epoch=E-step=S.ckpt`, keys: epoch, global_step, pytorch-lightning_version,
state_dict, callbacks, optimizer_states, lr_schedulers) so `render.py` / `NeRFNetwork.load_from_checkpoint` read them.

Data parallelism: one process per GPU (torchrun), identical replicas, every rank draws its own ray
batch, and ONE NCCL all-reduce per step over a flat fp32 buffer that all 40 parameter gradients are views of.
"""
import json
import os
import time
from pathlib import Path

import torch
import torch.distributed as dist


class JsonLogger:
    """Offline logger (wandb needs a network): one JSON line per logged step in <save_dir>/metrics.jsonl."""

    def __init__(self, name="run", project="NeRF", save_dir="."):
        self.name, self.project = name, project
        self.dir = Path(save_dir) / project / name
        self.dir.mkdir(parents=True, exist_ok=True)
        self.fh = open(self.dir / "metrics.jsonl", "a")

    def log_hyperparams(self, args):
        self.fh.write(json.dumps({"hyperparams": {k: str(v) for k, v in vars(args).items()}}) + "\n")
        self.fh.flush()

    def log_metrics(self, metrics, step):
        self.fh.write(json.dumps({"step": step, **metrics}) + "\n")
        self.fh.flush()

    def log_image(self, key, images, caption=None):
        from PIL import Image
        for i, im in enumerate(images):
            Image.fromarray(im).save(self.dir / f"{key}_{int(time.time())}_{i}.png")


class FlatGradients:
    """All parameter gradients as views of one contiguous fp32 buffer (one all-reduce per step)."""

    def __init__(self, params, optimizer=None):
        self.params = [p for p in params if p.requires_grad]
        if optimizer is not None and hasattr(optimizer, "flat_grads"):
            self.flat = optimizer.flat_grads          # optim.FlatAdam already keeps every .grad as a view of one buffer
            return
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())

    def norm(self):
        return self.flat.norm()


class Trainer:
    def __init__(self, gpus=0, default_root_dir=".", max_steps=100000, resume_from_checkpoint=None, logger=None,
                 check_val_every_n_epoch=10, track_grad_norm=-1, reload_dataloaders_every_n_epochs=0, callbacks=None,
                 log_every_n_steps=50, save_checkpoints=True, max_epochs=None):
        if not torch.cuda.is_available():
            raise RuntimeError("Trainer: this NeRF path has no CPU implementation; a CUDA device is required")
        self.root, self.max_steps, self.resume = Path(default_root_dir), max_steps, resume_from_checkpoint
        self.logger, self.val_every, self.track_grad_norm = logger, check_val_every_n_epoch, track_grad_norm
        self.reload_every = reload_dataloaders_every_n_epochs
        self.log_every, self.save_checkpoints, self.max_epochs = log_every_n_steps, save_checkpoints, max_epochs
        self.current_epoch, self.global_step = 0, 0
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._metrics = {}
        self.last_checkpoint = None

    def record(self, name, value):
        """Keeps the device tensor: converting here would synchronise the stream on every `self.log` call (8 per training
        step); the values are read back only when a line is actually written (`metrics()`)."""
        self._metrics[name] = value.detach() if torch.is_tensor(value) else value

    def metrics(self):
        return {k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else v) for k, v in self._metrics.items()
                if not torch.is_tensor(v) or v.numel() == 1}

    def checkpoint_dir(self):
        run = getattr(self.logger, "name", "run") if self.logger is not None else "run"
        return self.root / "NeRF" / str(run) / "checkpoints"

    def save_checkpoint(self, model, optimizer, scheduler):
        d = self.checkpoint_dir()
        d.mkdir(parents=True, exist_ok=True)
        path = d / f"epoch={self.current_epoch}-step={self.global_step - 1}.ckpt"
        torch.save({"epoch": self.current_epoch, "global_step": self.global_step, "pytorch-lightning_version": "1.5.10",
                    "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "callbacks": {},
                    "optimizer_states": [optimizer.state_dict()],
                    "lr_schedulers": [scheduler.state_dict()] if scheduler is not None else []}, str(path))
        self.last_checkpoint = path
        return path

    def fit(self, model, datamodule=None, train_dataloaders=None, val_dataloaders=None):
        model = model.to(self.device)
        model.trainer, model.logger = self, self.logger
        if datamodule is not None:
            datamodule.trainer = self
        cfg = model.configure_optimizers()
        optimizer = cfg["optimizer"] if isinstance(cfg, dict) else cfg
        scheduler = cfg.get("lr_scheduler") if isinstance(cfg, dict) else None
        if self.resume:
            ckpt = torch.load(str(self.resume), map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["state_dict"])
            if ckpt.get("optimizer_states"):
                optimizer.load_state_dict(ckpt["optimizer_states"][0])
            if scheduler is not None and ckpt.get("lr_schedulers"):
                scheduler.load_state_dict(ckpt["lr_schedulers"][0])
            self.current_epoch, self.global_step = ckpt.get("epoch", 0) + 1, ckpt.get("global_step", 0)
        grads = FlatGradients(model.parameters(), optimizer)
        loader = None
        while self.global_step < self.max_steps and (self.max_epochs is None or self.current_epoch < self.max_epochs):
            if loader is None or (self.reload_every and self.current_epoch % self.reload_every == 0):
                loader = datamodule.train_dataloader() if datamodule is not None else train_dataloaders
            model.train()
            for idx, batch in enumerate(loader):
                if self.global_step >= self.max_steps:
                    break
                batch = {k: v.to(self.device) for k, v in batch.items()}
                grads.zero()
                loss = model.training_step(batch, idx)
                loss.backward()
                grads.all_reduce_mean()
                if self.track_grad_norm and self.track_grad_norm > 0:
                    self.record("grad_2.0_norm_total", grads.norm())
                optimizer.step()
                self.global_step += 1
                if self.logger is not None and self.rank == 0 and self.global_step % self.log_every == 0:
                    self.logger.log_metrics(dict(self.metrics(), lr=optimizer.param_groups[0]["lr"], epoch=self.current_epoch),
                                            self.global_step)
            if scheduler is not None:
                scheduler.step()
            if self.val_every and (self.current_epoch + 1) % self.val_every == 0:
                vl = datamodule.val_dataloader() if datamodule is not None else val_dataloaders
                if vl is not None:
                    model.eval()
                    with torch.no_grad():
                        for idx, batch in enumerate(vl):
                            model.validation_step({k: v.to(self.device) for k, v in batch.items()}, idx)
            if self.save_checkpoints and self.rank == 0:
                self.save_checkpoint(model, optimizer, scheduler)
            self.current_epoch += 1
        return model
