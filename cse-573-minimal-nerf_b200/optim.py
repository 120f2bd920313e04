"""Adam for NeRFNetwork.configure_optimizers (nerf_model.py:134-143) as one hand-written kernel over flat buffers.

`FlatAdam` keeps every parameter, gradient and Adam moment of the model as views of four contiguous fp32 buffers and
updates all of them with a single `nerf_adam_step` launch (csrc/adam.cu) - the same arithmetic, in the same order, as
torch.optim.Adam's single-tensor path.  It is a `torch.optim.Optimizer`, so `ExponentialLR`, `state_dict()` /
`load_state_dict()` and the PL-format checkpoints (`optimizer_states`) keep working, with torch.optim.Adam's state layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter; param_groups with lr / betas / eps / weight_decay / amsgrad).

The kernel writes the parameters behind autograd's back.  So that nothing downstream keeps using stale derived data (the
packed bf16 weight images of nerf_model.NeRFModel are cached on (data_ptr, _version)), `step()` itself bumps every parameter's
version counter and then calls `on_params_changed` (NeRFNetwork.configure_optimizers passes its one-launch repack): an optimiser
built directly with `FlatAdam(net.parameters())` therefore still invalidates the caches.  `step()` also checks that the parameters
are still views of the flat buffer (a later `model.to()` / `.float()` would silently detach them from what the kernel updates).

Data parallel: `grad_scale` (default 1.0) multiplies every gradient as the kernel reads it; trainer.Trainer sets it to
1 / world_size, so the all-reduced SUM is never rescaled by a separate elementwise launch.
"""
import torch

import _native as nat


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, on_params_changed=None):
        params = [p for p in params if p.requires_grad]
        if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise RuntimeError("FlatAdam: expected fp32 CUDA parameters (this path has no CPU implementation)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._params = params
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self._n = (n + 3) // 4 * 4
        self.flat_params = torch.zeros(self._n, device=dev, dtype=torch.float32)
        self.flat_grads = torch.zeros(self._n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(self._n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(self._n, device=dev, dtype=torch.float32)
        self._step = 0
        self.grad_scale = 1.0
        self.on_params_changed = on_params_changed
        # graph-safe form (training.GraphedTrainStep): every step-dependent scalar lives in device memory
        self.graph_safe = False
        self.dev_state = torch.zeros(8, device=dev, dtype=torch.float32)      # lr, beta1, beta2, eps, grad_scale | 2 scratch
        self.dev_step = torch.zeros((), device=dev, dtype=torch.int64)
        self._dev_key = None
        off = 0
        for p in params:
            k = p.numel()
            self.flat_params[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + k].view_as(p)         # parameters become views of the flat buffer
            p.grad = self.flat_grads[off:off + k].view_as(p)           # and so do their gradients (one all-reduce per step)
            off += k
        self._bind_state()

    def _bind_state(self):
        """Per-parameter state entries in torch.optim.Adam's layout, as views of the flat moment buffers."""
        off = 0
        for p in self._params:
            k = p.numel()
            self.state[p] = {"step": torch.tensor(float(self._step)),
                             "exp_avg": self.flat_m[off:off + k].view_as(p),
                             "exp_avg_sq": self.flat_v[off:off + k].view_as(p)}
            off += k

    def zero_grad(self, set_to_none=False):
        self.flat_grads.zero_()

    def _grads_are_flat(self):
        off = 0
        base = self.flat_grads.data_ptr()
        for p in self._params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * off or not p.grad.is_contiguous():
                return False
            off += p.numel()
        return True

    def _params_are_flat(self):
        off = 0
        base = self.flat_params.data_ptr()
        for p in self._params:
            if p.data_ptr() != base + 4 * off:
                return False
            off += p.numel()
        return True

    def params_changed(self):
        """Call after writing the parameters through the flat buffer (the step kernel, a broadcast into `flat_params`, ...):
        bumps the version counters autograd-side caches are keyed on, then runs the `on_params_changed` callback."""
        torch.autograd.graph.increment_version(self._params)
        if self.on_params_changed is not None:
            self.on_params_changed()

    def sync_device_state(self):
        """Graph-safe form: push lr / betas / eps / grad_scale and the step count to the device when the host's view changed (the
        per-epoch LR decay, a loaded checkpoint).  Host -> device copies: call it OUTSIDE graph capture."""
        g = self.param_groups[0]
        key = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(self.grad_scale))
        if key != self._dev_key:
            self.dev_state[:5].copy_(torch.tensor(key, dtype=torch.float32))
            self._dev_key = key
        return self

    def set_device_step(self):
        self.dev_step.fill_(self._step)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if not self._params_are_flat():
            raise RuntimeError("FlatAdam: a parameter no longer points into the optimiser's flat buffer (was the model moved or "
                               "cast after configure_optimizers()?); rebuild the optimiser after model.to() / .float()")
        if not self._grads_are_flat():            # someone re-pointed .grad (e.g. trainer.FlatGradients): gather once
            off = 0
            for p in self._params:
                k = p.numel()
                if p.grad is not None:
                    self.flat_grads[off:off + k].copy_(p.grad.reshape(-1))
                else:
                    self.flat_grads[off:off + k].zero_()
                off += k
        g = self.param_groups[0]
        self._step += 1
        if self.graph_safe:
            if not torch.cuda.is_current_stream_capturing():
                self.sync_device_state()            # an eager step after graphed ones: the scheduler may have moved lr
            with torch.cuda.device(self.flat_params.device):
                nat.check(nat.lib().nerf_adam_step_dev(nat.ptr(self.flat_params), nat.ptr(self.flat_grads), nat.ptr(self.flat_m),
                                                       nat.ptr(self.flat_v), self._n, nat.ptr(self.dev_state), nat.ptr(self.dev_step),
                                                       nat.stream()), "nerf_adam_step_dev")
                self.params_changed()
            return loss
        with torch.cuda.device(self.flat_params.device):
            nat.check(nat.lib().nerf_adam_step(nat.ptr(self.flat_params), nat.ptr(self.flat_grads), nat.ptr(self.flat_m),
                                               nat.ptr(self.flat_v), self._n, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                               float(g["eps"]), self._step, float(self.grad_scale), nat.stream()), "nerf_adam_step")
            self.params_changed()
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout; the per-parameter `step` entries are refreshed here instead of 40 times per step."""
        for p in self._params:
            self.state[p]["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)        # torch copies the saved tensors in; move them back into the flat buffers
        off = 0
        for p in self._params:
            k = p.numel()
            st = self.state.get(p, {})
            if "exp_avg" in st:
                self.flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1).to(self.flat_m.device))
                self.flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1).to(self.flat_v.device))
                self._step = int(float(st.get("step", self._step)))
            off += k
        self._bind_state()
        self.set_device_step()                     # graph-safe form: the device copies follow the loaded state
        self._dev_key = None
