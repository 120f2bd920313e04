"""Minimal stand-in for the slice of pytorch_lightning.Trainer that the reference's train_nerf.py uses
(train_nerf.py:20-34): max_steps, per-epoch LR scheduler step, resume_from_checkpoint, validation every n epochs,
dataloader reload (the cropping switch), gradient-norm tracking, and PL-1.5.10-shaped checkpoints
(`default_root_dir/NeRF/<run>/checkpoints/epoch=E-step=S.ckpt`, keys: epoch, global_step, pytorch-lightning_version,
state_dict, callbacks, optimizer_states, lr_schedulers) so `render.py` / `NeRFNetwork.load_from_checkpoint` read them.

Data parallelism: one process per GPU (torchrun).  Before the first step rank 0's parameters (and, on a
resume, its Adam moments and step count) are broadcast so that the replicas are IDENTICAL, then every rank seeds its data
sampling with `seed + rank` so that the ranks draw DIFFERENT images / pixels on purpose.  Per step the gradients of all ranks are
summed by NCCL over the flat fp32 buffer all 40 `.grad`s are views of - the coarse network's slice as soon as its backward has been
queued (it overlaps the fine network's backward), the rest after the backward - and the 1 / world_size factor is applied inside
the Adam kernel (`FlatAdam.grad_scale`), so every rank takes the same step from the same parameters.  Validation and logging run
on rank 0 only, with the frame rendered locally (no collective depends on per-rank control flow).
"""
import json
import time
from pathlib import Path

import torch
import torch.distributed as dist


class JsonLogger:
    """Offline logger (wandb needs a network): one JSON line per logged step in <save_dir>/metrics.jsonl.
    `enabled=False` (every rank but 0 under torchrun) keeps the name - checkpoint paths derive from it - and writes nothing."""

    def __init__(self, name="run", project="NeRF", save_dir=".", enabled=True):
        self.name, self.project, self.enabled = name, project, enabled
        self.dir = Path(save_dir) / project / name
        self.fh = None
        if enabled:
            self.dir.mkdir(parents=True, exist_ok=True)
            self.fh = open(self.dir / "metrics.jsonl", "a")

    def log_hyperparams(self, args):
        if self.fh is not None:
            self.fh.write(json.dumps({"hyperparams": {k: str(v) for k, v in vars(args).items()}}) + "\n")
            self.fh.flush()

    def log_metrics(self, metrics, step):
        if self.fh is not None:
            self.fh.write(json.dumps({"step": step, **metrics}) + "\n")
            self.fh.flush()

    def log_image(self, key, images, caption=None):
        if self.fh is None:
            return
        from PIL import Image
        for i, im in enumerate(images):
            Image.fromarray(im).save(self.dir / f"{key}_{int(time.time())}_{i}.png")


class FlatGradients:
    """All parameter gradients as views of one contiguous fp32 buffer, reduced across the data-parallel ranks once per step.

    `reduce_async(lo, hi)` queues the SUM of a slice whose gradients are final (NCCL runs it on its own stream, ordered behind
    what the current stream has queued so far); `all_reduce_mean()` reduces whatever has not been queued yet, waits for the
    pending slices and applies 1 / world_size - inside the optimiser's kernel when it has a `grad_scale` (optim.FlatAdam), by one
    `mul_` otherwise."""

    def __init__(self, params, optimizer=None):
        self.params = [p for p in params if p.requires_grad]
        self.optimizer = optimizer
        self._pending, self._done_upto = [], 0
        if optimizer is not None and hasattr(optimizer, "flat_grads"):
            self.flat = optimizer.flat_grads          # optim.FlatAdam already keeps every .grad as a view of one buffer
        else:
            n = sum(p.numel() for p in self.params)
            self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
            off = 0
            for p in self.params:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.folded = optimizer is not None and hasattr(optimizer, "grad_scale")
        if self.folded:
            optimizer.grad_scale = 1.0 / self.world

    def zero(self):
        self.flat.zero_()
        self._pending, self._done_upto = [], 0

    def reduce_async(self, hi):
        """Gradients [done, hi) of the flat buffer are final: start summing them over the ranks."""
        if self.world > 1 and hi > self._done_upto:
            self._pending.append(dist.all_reduce(self.flat[self._done_upto:hi], op=dist.ReduceOp.SUM, async_op=True))
            self._done_upto = hi

    def all_reduce_mean(self):
        if self.world > 1:
            self.reduce_async(self.flat.numel())
            for work in self._pending:
                work.wait()
            self._pending = []
            if not self.folded:
                self.flat.mul_(1.0 / self.world)

    def norm(self):
        """2-norm of the MEAN gradient (what the optimiser applies)."""
        n = self.flat.norm()
        return n / self.world if (self.folded and self.world > 1) else n


class Trainer:
    def __init__(self, gpus=0, default_root_dir=".", max_steps=100000, resume_from_checkpoint=None, logger=None,
                 check_val_every_n_epoch=10, track_grad_norm=-1, reload_dataloaders_every_n_epochs=0, callbacks=None,
                 log_every_n_steps=50, save_checkpoints=True, max_epochs=None, seed=None, cuda_graph=True):
        if not torch.cuda.is_available():
            raise RuntimeError("Trainer: this NeRF path has no CPU implementation; a CUDA device is required")
        self.root, self.max_steps, self.resume = Path(default_root_dir), max_steps, resume_from_checkpoint
        self.logger, self.val_every, self.track_grad_norm = logger, check_val_every_n_epoch, track_grad_norm
        self.reload_every = reload_dataloaders_every_n_epochs
        self.log_every, self.save_checkpoints, self.max_epochs = log_every_n_steps, save_checkpoints, max_epochs
        self.current_epoch, self.global_step = 0, 0
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.seed = seed
        self.cuda_graph = cuda_graph            # capture the training step as a CUDA graph where the data allows it (training.GraphedTrainStep)
        self._steppers = {}
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
        previous, self.last_checkpoint = self.last_checkpoint, path
        if previous is not None and previous != path and Path(previous).exists():
            Path(previous).unlink()            # Lightning's default ModelCheckpoint keeps the latest file only (save_top_k = 1)
        return path

    def fit(self, model, datamodule=None, train_dataloaders=None, val_dataloaders=None):
        if hasattr(model, "check_trainable"):
            model.check_trainable()                        # e.g. train_nerf.py -p / -d other than 10 / 4: fail before any work
        model = model.to(self.device)
        model.trainer, model.logger = self, self.logger
        if datamodule is not None:
            datamodule.trainer = self
        cfg = model.configure_optimizers()
        optimizer = cfg["optimizer"] if isinstance(cfg, dict) else cfg
        scheduler = cfg.get("lr_scheduler") if isinstance(cfg, dict) else None
        self.optimizer, self.scheduler = optimizer, scheduler
        if self.resume:
            ckpt = torch.load(str(self.resume), map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["state_dict"])
            if ckpt.get("optimizer_states"):
                optimizer.load_state_dict(ckpt["optimizer_states"][0])
            if scheduler is not None and ckpt.get("lr_schedulers"):
                scheduler.load_state_dict(ckpt["lr_schedulers"][0])
            self.current_epoch, self.global_step = ckpt.get("epoch", 0) + 1, ckpt.get("global_step", 0)
        grads = FlatGradients(model.parameters(), optimizer)
        self.synchronize_replicas(model, optimizer)
        n_coarse = sum(p.numel() for p in model.coarse_network.parameters()) if hasattr(model, "coarse_network") else 0
        if self.world > 1 and n_coarse and hasattr(optimizer, "flat_grads"):
            # the coarse network's gradients are final before the fine network's backward starts (training.RenderFunction):
            # their share of the all-reduce overlaps it
            model.on_coarse_grads_ready = lambda: grads.reduce_async(n_coarse)
        loader = None
        while self.global_step < self.max_steps and (self.max_epochs is None or self.current_epoch < self.max_epochs):
            if loader is None or (self.reload_every and self.current_epoch % self.reload_every == 0):
                loader = datamodule.train_dataloader() if datamodule is not None else train_dataloaders
            model.train()
            stepper = self._graphed_stepper(model, optimizer, grads, loader)
            if stepper is not None:
                self._run_epoch_graphed(stepper, loader, optimizer)
            else:
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
                    self._step_done(optimizer)
            if scheduler is not None:
                scheduler.step()
            if self.val_every and (self.current_epoch + 1) % self.val_every == 0:
                self.validate(model, datamodule.val_dataloader() if datamodule is not None else val_dataloaders)
            if self.save_checkpoints and self.rank == 0:
                self.save_checkpoint(model, optimizer, scheduler)
            self.current_epoch += 1
        model.on_coarse_grads_ready = None
        # the captured training steps hold NCCL work (data parallel) and a few GB of graph-private memory: release them with the
        # run, before anybody tears the process group down
        torch.cuda.synchronize(self.device)
        for stepper in self._steppers.values():
            stepper.close()
        self._steppers.clear()
        return model

    def _step_done(self, optimizer):
        self.global_step += 1
        if self.logger is not None and self.rank == 0 and self.global_step % self.log_every == 0:
            self.logger.log_metrics(dict(self.metrics(), lr=optimizer.param_groups[0]["lr"], epoch=self.current_epoch), self.global_step)

    def _graphed_stepper(self, model, optimizer, grads, loader):
        """A training.GraphedTrainStep for this loader's dataset, or None when the step has to run eagerly: the loader is not the
        device-resident Blender-synthetic one, the optimiser is not optim.FlatAdam, or too few steps are left to pay for the capture."""
        ds = getattr(loader, "dataset", None)
        if (not self.cuda_graph or ds is None or not hasattr(ds, "stacked") or not hasattr(optimizer, "graph_safe")
                or not hasattr(model, "coarse_network") or self.max_steps - self.global_step < 8 or len(ds) < 1):
            return None
        key = (id(ds), bool(ds.cropping))
        if key not in self._steppers:
            import training
            images, poses = ds.stacked()
            order = self._epoch_order(loader)
            self._pending_order = order
            self._steppers[key] = training.GraphedTrainStep(model, optimizer, grads, images, poses, ds.focal, ds.num_rays, cropping=ds.cropping,
                                                            track_grad_norm=bool(self.track_grad_norm and self.track_grad_norm > 0),
                                                            warmup_images=order[:3])
            for _ in range(3):                             # the warm-up steps before the capture were real optimiser steps
                self._step_done(optimizer)
            self._pending_order = order[3:]
        return self._steppers[key]

    @staticmethod
    def _epoch_order(loader):
        n = len(loader.dataset)
        return torch.randperm(n).tolist() if getattr(loader, "shuffle", False) else list(range(n))

    def _run_epoch_graphed(self, stepper, loader, optimizer):
        order = getattr(self, "_pending_order", None)
        self._pending_order = None
        if order is None:
            order = self._epoch_order(loader)
        for image_index in order:
            if self.global_step >= self.max_steps:
                break
            stepper.step(image_index)
            if stepper.grad_norm is not None:
                self.record("grad_2.0_norm_total", stepper.grad_norm)
            self._step_done(optimizer)

    def synchronize_replicas(self, model, optimizer):
        """Data parallel: rank 0's parameters, Adam moments and step count become everybody's; then every rank seeds its own
        data sampling (torch CPU + CUDA generators, Python's `random`) with seed + rank.  The base seed is the `seed` argument, or
        drawn by rank 0 and broadcast.  A single process only seeds when `seed` was given."""
        import random
        if self.world > 1:
            flat = [getattr(optimizer, n) for n in ("flat_params", "flat_m", "flat_v") if hasattr(optimizer, n)]
            if flat:
                for t in flat:
                    dist.broadcast(t, src=0)
                step = torch.tensor([getattr(optimizer, "_step", 0)], device=self.device, dtype=torch.int64)
                dist.broadcast(step, src=0)
                optimizer._step = int(step.item())
                optimizer.params_changed()                 # version bumps + re-pack of the bf16 weight images
            else:                                          # any other optimiser: parameters one by one
                for p in model.parameters():
                    dist.broadcast(p.data, src=0)
                if hasattr(model, "invalidate_packed_weights"):
                    model.invalidate_packed_weights()
            base = torch.tensor([self.seed if self.seed is not None else torch.seed() % (2 ** 31)], device=self.device, dtype=torch.int64)
            dist.broadcast(base, src=0)
            self.seed = int(base.item())
        if self.seed is not None:
            torch.manual_seed(self.seed + self.rank)
            random.seed(self.seed + self.rank)

    def validate(self, model, loader):
        """The reference's every-n-epochs validation (train_nerf.py:27, nerf_model.py:171-205) on rank 0 only; the frame it logs
        is rendered by this rank alone (multi_gpu.local_only), so no collective sits behind per-rank control flow.  The other
        ranks wait at the barrier."""
        import multi_gpu
        if loader is not None and self.rank == 0:
            model.eval()
            with torch.no_grad(), multi_gpu.local_only():
                for idx, batch in enumerate(loader):
                    model.validation_step({k: v.to(self.device) for k, v in batch.items()}, idx)
            model.train()
        if self.world > 1:
            dist.barrier()
