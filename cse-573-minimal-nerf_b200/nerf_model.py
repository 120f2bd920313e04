"""B200 drop-in for the reference's `nerf_model.py`: same classes, constructor arguments, state_dict keys,
forward signatures and return shapes; the arithmetic runs in libnerf_b200.so.

Reference counterparts (file:line): positional_encoding 19-33, normalize_coordinates 35-54,
NeRFNetwork 56-205, NeRFModel 308-389.

`precision`:
  "bf16" (default) - the fused tcgen05 kernel: bf16 operands, fp32 accumulation in TMEM (position_dim 10,
                     direction_dim 4 only);
  "fp32"           - the exact-fp32 CUDA-core kernel (any encoding sizes); tight-parity mode.
"""
import ctypes
import math
import random
from timeit import default_timer as timer

import torch
import torch.nn as nn
import torch.nn.functional as F

import _native as nat
import nerf_helpers
from lightning_shim import LightningModule

ACT_FN = nn.ReLU()


@nat.device_guard
def positional_encoding(x, dim=10):
    """[..., C] -> [..., 2*dim*C]: per frequency i, cos(2^i pi x) for all channels then sin(2^i pi x)."""
    t = nat.dev(x, "x")
    c = t.shape[-1]
    flat = t.reshape(-1, c)
    out = torch.empty((flat.shape[0], 2 * dim * c), device=t.device, dtype=torch.float32)
    nat.check(nat.lib().nerf_positional_encoding(nat.ptr(flat), flat.shape[0], c, dim, nat.ptr(out), nat.stream()),
              "nerf_positional_encoding")
    return out.reshape(*t.shape[:-1], 2 * dim * c)


def normalize_coordinates(x, bound=math.pi):
    """x / bound (coordinates are expected within [-bound, bound])."""
    return x / bound


class NeRFModel(nn.Module):
    """One NeRF MLP: PE(x) -> 4x256 -> [+PE(x)] -> 3x256 -> sigma (ReLU) and rgb (128 -> 3, sigmoid).
    The nn.Linear modules only hold the parameters (identical state_dict keys to the reference)."""

    def __init__(self, position_dim=10, direction_dim=4, precision="bf16"):
        super().__init__()
        self.position_dim, self.direction_dim, self.precision = position_dim, direction_dim, precision
        pe, de = position_dim * 2 * 3, direction_dim * 2 * 3
        self.mlp = nn.Sequential(nn.Linear(pe, 256), ACT_FN, nn.Linear(256, 256), ACT_FN,
                                 nn.Linear(256, 256), ACT_FN, nn.Linear(256, 256), ACT_FN)
        self.feature_fn = nn.Sequential(nn.Linear(256 + pe, 256), ACT_FN, nn.Linear(256, 256), ACT_FN,
                                        nn.Linear(256, 256))
        self.density_fn = nn.Sequential(nn.Linear(256, 1), nn.ReLU())
        self.rgb_fn = nn.Sequential(nn.Linear(256 + de, 128), ACT_FN, nn.Linear(128, 3), nn.Sigmoid())
        self._packed = None
        self._packed_key = None
        self._packed_t = None
        self._packed_t_key = None

    # ---- parameter plumbing
    def ordered_params(self):
        """The 20 tensors in state_dict order."""
        mods = [self.mlp[0], self.mlp[2], self.mlp[4], self.mlp[6], self.feature_fn[0], self.feature_fn[2],
                self.feature_fn[4], self.density_fn[0], self.rgb_fn[0], self.rgb_fn[2]]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return out

    def _param_ptrs(self):
        ps = [nat.dev(p.detach(), "parameter") for p in self.ordered_params()]
        arr = (ctypes.c_void_p * 20)(*[p.data_ptr() for p in ps])
        return arr, ps

    def uses_tensor_cores(self):
        return self.precision == "bf16" and self.position_dim == 10 and self.direction_dim == 4

    @nat.device_guard
    def packed_weights(self):
        """bf16 swizzled weight image for the tcgen05 kernel; re-packed whenever a parameter changed."""
        params = self.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or self._packed_key != key or self._packed.device != params[0].device:
            arr, keep = self._param_ptrs()
            if self._packed is None or self._packed.device != params[0].device:
                self._packed = torch.empty(nat.lib().nerf_packed_bytes(), dtype=torch.uint8, device=params[0].device)
            nat.check(nat.lib().nerf_pack_weights(arr, nat.ptr(self._packed), nat.stream()), "nerf_pack_weights")
            self._packed_key = key
        return self._packed

    @nat.device_guard
    def packed_weights_t(self):
        """W^T stage image for the tcgen05 dgrad kernel (training); re-packed whenever a parameter changed."""
        params = self.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed_t is None or self._packed_t_key != key or self._packed_t.device != params[0].device:
            arr, keep = self._param_ptrs()
            if self._packed_t is None or self._packed_t.device != params[0].device:
                self._packed_t = torch.empty(nat.lib().nerf_packed_t_bytes(), dtype=torch.uint8, device=params[0].device)
            nat.check(nat.lib().nerf_pack_weights_t(arr, nat.ptr(self._packed_t), nat.stream()), "nerf_pack_weights_t")
            self._packed_t_key = key
        return self._packed_t

    # ---- forward
    @nat.device_guard
    def forward(self, samples, direc):
        """samples [N,S,3], direc [N,3] -> density [N,S,1], rgb [N,S,3]."""
        x, dr = nat.dev(samples, "samples"), nat.dev(direc, "direc")
        N, S, _ = x.shape
        sigma = torch.empty((N, S, 1), device=x.device, dtype=torch.float32)
        rgb = torch.empty((N, S, 3), device=x.device, dtype=torch.float32)
        if self.uses_tensor_cores():
            nat.check(nat.lib().nerf_mlp_forward_tc_points(nat.ptr(self.packed_weights()), nat.ptr(x), nat.ptr(dr), N, S,
                                                           nat.ptr(sigma), nat.ptr(rgb), nat.stream()), "nerf_mlp_forward_tc")
        else:
            arr, keep = self._param_ptrs()
            nat.check(nat.lib().nerf_mlp_forward_fp32(arr, self.position_dim, self.direction_dim, nat.ptr(x), nat.ptr(dr),
                                                      N, S, nat.ptr(sigma), nat.ptr(rgb), nat.stream()), "nerf_mlp_forward_fp32")
        return sigma, rgb

    @nat.device_guard
    def forward_rays(self, o_rays, d_rays, ts):
        """Same network evaluated at o + t*d for ts [N,S,1] without materialising the points."""
        N, S = ts.shape[0], ts.shape[1]
        sigma = torch.empty((N, S, 1), device=ts.device, dtype=torch.float32)
        rgb = torch.empty((N, S, 3), device=ts.device, dtype=torch.float32)
        if self.uses_tensor_cores():
            packed = self.packed_weights()
            with nat.timed_kernel("mlp_tc_kernel", N * S):
                nat.check(nat.lib().nerf_mlp_forward_tc(nat.ptr(packed), nat.ptr(o_rays), nat.ptr(d_rays), nat.ptr(ts),
                                                        N, S, nat.ptr(sigma), nat.ptr(rgb), nat.stream()), "nerf_mlp_forward_tc")
            return sigma, rgb
        pts = d_rays[:, None, :] * ts + o_rays[:, None, :]
        return self.forward(pts.contiguous(), d_rays)


    def can_composite(self, S):
        """True when render_rays can run (tensor-core shape, sample count the fused kernel supports)."""
        return self.uses_tensor_cores() and bool(nat.lib().nerf_mlp_composite_tc_supported(int(S)))

    @nat.device_guard
    def render_rays(self, o_rays, d_rays, ts, want_weights=True, keep_samples=False, save=False, stats=None, out=None, strata=None):
        """Network + alpha compositing in ONE kernel (nerf_mlp_composite_tc): the per-sample sigma / rgb stay on the SM
        unless `keep_samples` (or `save`, the training form, which also stores activations + ReLU sign words).
        Returns the same dict as nerf_helpers.composite plus 'sigma', 'rgb_samples', 'saved' (None when not kept).
        `stats`: optional ZEROED [4] fp32 buffer for the density statistics (one is allocated otherwise); `out`: optional
        contiguous [N,3] fp32 tensor the ray colours are written into (e.g. a slice of the frame buffer).
        `strata` = (u [N,S], t_base [S], step) with ts=None: the stratified depths of generate_coarse_samples are formed inside
        the kernel as well (returned as 'ts', bit-identical to nerf_coarse_sample)."""
        if strata is not None:
            u, t_base, step = strata
            u = nat.dev(u, "u")
            N, S = u.shape[0], u.shape[1]
            dv = u.device
            ts = torch.empty((N, S, 1), device=dv, dtype=torch.float32)
        else:
            N, S = ts.shape[0], ts.shape[1]
            dv = ts.device
        keep = keep_samples or save
        sigma = torch.empty((N, S, 1), device=dv, dtype=torch.float32) if keep else None
        rgb = torch.empty((N, S, 3), device=dv, dtype=torch.float32) if keep else None
        acts = masks = None
        if save:
            import training
            rows = training.padded_rows(N * S)
            acts = torch.empty((rows * training.ACT,), device=dv, dtype=torch.bfloat16)
            masks = torch.empty((rows * (training.ACT // 64),), device=dv, dtype=torch.int64)
        w = torch.empty((N, S, 1), device=dv, dtype=torch.float32) if want_weights else None
        if out is not None and (out.shape != (N, 3) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dv):
            raise ValueError("render_rays: `out` must be a contiguous [N,3] fp32 tensor on the rays' device")
        col = out if out is not None else torch.empty((N, 3), device=dv, dtype=torch.float32)
        depth = torch.empty((N,), device=dv, dtype=torch.float32)
        acc = torch.empty((N,), device=dv, dtype=torch.float32)
        if stats is None:
            stats = torch.zeros((4,), device=dv, dtype=torch.float32)
        packed = self.packed_weights()
        with nat.timed_kernel("mlp_tc_kernel(train)" if save else "mlp_tc_kernel", N * S):
            if strata is not None:
                nat.check(nat.lib().nerf_mlp_composite_tc_strata(nat.ptr(packed), nat.ptr(o_rays), nat.ptr(d_rays), nat.ptr(u), nat.ptr(t_base),
                                                                 float(step), N, S, nat.ptr(ts), nat.ptr(sigma), nat.ptr(rgb), nat.ptr(acts),
                                                                 nat.ptr(masks), nat.ptr(w), nat.ptr(col), nat.ptr(depth), nat.ptr(acc),
                                                                 nat.ptr(stats), nat.stream()), "nerf_mlp_composite_tc_strata")
            else:
                nat.check(nat.lib().nerf_mlp_composite_tc(nat.ptr(packed), nat.ptr(o_rays), nat.ptr(d_rays), nat.ptr(ts), N, S,
                                                          nat.ptr(sigma), nat.ptr(rgb), nat.ptr(acts), nat.ptr(masks),
                                                          nat.ptr(w), nat.ptr(col), nat.ptr(depth), nat.ptr(acc), nat.ptr(stats),
                                                          nat.stream()), "nerf_mlp_composite_tc")
        return {"weights": w, "rgb": col, "depth": depth, "acc": acc, "stats": stats[:2], "norm": stats[2],
                "sigma": sigma, "rgb_samples": rgb, "saved": (acts, masks) if save else None, "ts": ts}


class NeRFNetwork(LightningModule):
    """Coarse + fine NeRF (the reference's Lightning module): forward(o_rays, d_rays) ->
    {'fine_rgb_rays': [N,3], 'coarse_rgb_rays': [N,3]}."""

    def __init__(self, position_dim=10, direction_dim=4, coarse_samples=64,
                 fine_samples=128, near=2.0, far=6.0, precision="bf16"):
        super().__init__()
        self.position_dim, self.direction_dim = position_dim, direction_dim
        self.coarse_samples, self.fine_samples = coarse_samples, fine_samples
        self.near, self.far = near, far
        self.coarse_network = NeRFModel(position_dim, direction_dim, precision)
        self.fine_network = NeRFModel(position_dim, direction_dim, precision)
        self.im_idx, self.max_idx = 0, 1
        self.timer = timer()
        self.last = {}                      # depth / acc / weights of the most recent forward
        self.keep_samples = False           # True: inference passes also materialise per-sample sigma / rgb in self.last
        self.on_coarse_grads_ready = None   # trainer hook: called by the backward once the coarse network's gradients are queued

    @nat.device_guard
    def forward(self, o_rays, d_rays, rand=None, fine_out=None):
        """rand = (u_c [N,C], eps [N,1], u_f [N,F,1]) replaces the three torch.rand draws when given.
        fine_out (inference only): contiguous [N,3] fp32 tensor that receives 'fine_rgb_rays' directly (a slice of a frame
        buffer: saves the copy of view_reconstruction's chunk loop).
        With gradients enabled the result is differentiable w.r.t. the 40 parameters (training.RenderFunction)."""
        import training
        o, d = nat.dev(o_rays, "o_rays"), nat.dev(d_rays, "d_rays")
        params = list(self.parameters())
        # the differentiable path exists for the fused bf16 tensor-core kernels only; precision="fp32" networks
        # (tight-parity / odd encoding sizes) always run the inference path and return tensors without a grad_fn
        if torch.is_grad_enabled() and any(p.requires_grad for p in params) and self.coarse_network.uses_tensor_cores():
            u_c, eps, u_f = rand if rand is not None else (None, None, None)
            ordered = self.coarse_network.ordered_params() + self.fine_network.ordered_params()
            c_rgb, f_rgb = training.RenderFunction.apply(self, o, d, u_c, eps, u_f, *ordered)
        else:
            c_rgb, f_rgb, aux = training.forward_pass(self, o, d, rand, save=False, keep_samples=self.keep_samples, fine_out=fine_out)
            self._publish(aux)
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                # forward-only kernels: usable as before, but a backward through the result fails with the reason
                c_rgb, f_rgb = training.NotDifferentiable.apply(self.TRAINABLE_ONLY, params[0], c_rgb, f_rgb)
        return {'fine_rgb_rays': f_rgb, 'coarse_rgb_rays': c_rgb}

    TRAINABLE_ONLY = ("NeRFNetwork: the differentiable path exists for precision='bf16' with position_dim=10, direction_dim=4 (the fused "
                      "tcgen05 kernels) only; this network runs the forward-only exact-fp32 kernel - it can render, not train")

    def check_trainable(self):
        """Raises, at construction time of a training run, what a later loss.backward() would raise."""
        if not (self.coarse_network.uses_tensor_cores() and self.fine_network.uses_tensor_cores()):
            raise RuntimeError(self.TRAINABLE_ONLY + f" (got position_dim={self.position_dim}, direction_dim={self.direction_dim}, "
                               f"precision={self.coarse_network.precision!r})")

    def _publish(self, aux):
        """The four density statistics the reference logs inside forward (nerf_model.py:105-106,124-125) + extras."""
        c, f = aux["c"], aux["f"]
        self.log('coarse_density_norms', c["norm"], batch_size=1)
        self.log('coarse_density_non_zeros', c["stats"][1], batch_size=1)
        self.log('fine_density_norms', f["norm"], batch_size=1)
        self.log('fine_density_non_zeros', f["stats"][1], batch_size=1)
        self.last = {"depth": f["depth"], "acc": f["acc"], "ts": aux["ts"], "coarse_ts": aux["c_ts"],
                     "coarse_weights": c["weights"], "coarse_sigma": aux["c_sigma"], "fine_sigma": aux["f_sigma"],
                     "coarse_rgb": aux["c_rgb"], "fine_rgb": aux["f_rgb"]}

    @nat.device_guard
    def repack_all(self):
        """All four bf16 weight images (forward + W^T, both networks) in one launch; called after every optimiser step, whose
        in-place update does not bump the parameters' version counters (the keys the cached images are checked against)."""
        nets = (self.coarse_network, self.fine_network)
        if not all(n.uses_tensor_cores() for n in nets):
            return self.invalidate_packed_weights()
        params = nets[0].ordered_params() + nets[1].ordered_params()
        keep = [nat.dev(p.detach(), "parameter") for p in params]
        arr = (ctypes.c_void_p * 40)(*[p.data_ptr() for p in keep])
        dev = params[0].device
        for n in nets:
            if n._packed is None or n._packed.device != dev:
                n._packed = torch.empty(nat.lib().nerf_packed_bytes(), dtype=torch.uint8, device=dev)
            if n._packed_t is None or n._packed_t.device != dev:
                n._packed_t = torch.empty(nat.lib().nerf_packed_t_bytes(), dtype=torch.uint8, device=dev)
        nat.check(nat.lib().nerf_pack_weights_all(arr, nat.ptr(nets[0]._packed), nat.ptr(nets[0]._packed_t), nat.ptr(nets[1]._packed),
                                                  nat.ptr(nets[1]._packed_t), nat.stream()), "nerf_pack_weights_all")
        for n in nets:
            key = tuple((p.data_ptr(), p._version) for p in n.ordered_params())
            n._packed_key = n._packed_t_key = key

    def invalidate_packed_weights(self):
        """Force a re-pack of the bf16 weight images on the next forward / backward (call after changing parameters
        through anything that does not bump tensor versions)."""
        for net in (self.coarse_network, self.fine_network):
            net._packed_key = None
            net._packed_t_key = None

    def _coarse_ts(self, o, d, u_c):
        N, C = u_c.shape
        t_base, step = nerf_helpers._strata(self.near, self.far, C, o.device)
        ts = torch.empty((N, C, 1), device=o.device, dtype=torch.float32)
        nat.check(nat.lib().nerf_coarse_sample(nat.ptr(o), nat.ptr(d), nat.ptr(nat.dev(u_c, "u_c")), nat.ptr(t_base), step,
                                               N, C, None, nat.ptr(ts), nat.stream()), "nerf_coarse_sample")
        return ts

    def configure_optimizers(self):
        start_lr, end_lr, num_epochs = 5e-4, 5e-5, 1200
        gamma = (end_lr / start_lr) ** (1 / num_epochs)
        import optim
        # one hand-written kernel over flat parameter / gradient / moment buffers (csrc/adam.cu) instead of torch's
        # multi-tensor Adam.  After every update the optimiser bumps the parameters' version counters (what the packed bf16
        # weight images are keyed on) and calls back: all four images are re-packed by one launch
        optimizer = optim.FlatAdam(self.parameters(), lr=start_lr, on_params_changed=self.repack_all)
        lr_decay = torch.optim.lr_scheduler.ExponentialLR(optimizer=optimizer, gamma=gamma)
        return {'optimizer': optimizer, 'lr_scheduler': lr_decay}

    def _step(self, batch, prefix):
        nerf_helpers.fix_batchify(batch)
        o_rays, d_rays, rgb = batch['origin'], batch['direc'], batch['rgb']
        pred = self.forward(o_rays, d_rays)
        N = pred['fine_rgb_rays'].shape[0]
        coarse_loss = F.mse_loss(pred['coarse_rgb_rays'], rgb)
        fine_loss = F.mse_loss(pred['fine_rgb_rays'], rgb)
        loss = coarse_loss + fine_loss
        self.log(f'{prefix}_loss', loss, batch_size=N)
        self.log(f'{prefix}_fine_loss', fine_loss, batch_size=N)
        self.log(f'{prefix}_coarse_loss', coarse_loss, batch_size=N)
        return loss, N

    def training_step(self, train_batch, batch_idx):
        loss, N = self._step(train_batch, 'train')
        self.log('train iteration speed', timer() - self.timer, batch_size=N)
        self.timer = timer()
        return loss

    def validation_step(self, val_batch, batch_idx):
        self.max_idx = max(self.max_idx, batch_idx)
        if batch_idx == 0:
            self.im_idx = random.randint(0, self.max_idx)
        loss, N = self._step(val_batch, 'val')
        if batch_idx == self.im_idx and 'all_origin' in val_batch:
            im = nerf_helpers.view_reconstruction(self, val_batch['all_origin'], val_batch['all_direc'], N=N)
            if self.logger is not None and hasattr(self.logger, 'log_image'):
                self.logger.log_image(key='recon', images=[im], caption=[f'val/{self.im_idx}.png'])
        return loss
