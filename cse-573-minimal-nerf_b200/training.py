"""Training path of NeRFNetwork: fused forward that keeps activations + the backward chain.

`RenderFunction` is the autograd node behind `NeRFNetwork.forward` when gradients are enabled.  Forward =
the same kernels as inference (stratified sampling, fused tcgen05 MLP, compositing, inverse-CDF sampling, merge
sort) with the MLP launched in its training form, which also stores the bf16 activations every layer consumed
([samples, 1920] per network).  Backward =
  * hand-written compositing backward (`nerf_composite_backward`): dL/d ray colour -> dL/d(sigma, rgb) pre-activations
  * the dgrad / wgrad chain through the 10 Linear layers of each network as hand-written tcgen05 kernels
    (`nerf_mlp_backward_tc`, `nerf_wgrad_tc`).  (The same chain as bf16 cuBLAS GEMMs, the on-device reference those kernels are
    validated against, lives with the tests: tests/library_backward.py.)
No gradient flows from the fine loss into the coarse network: the sampler's indices and depths carry none
(nerf_model.py:114-120), so the two chains are independent.
"""
import torch

import _native as nat
import nerf_helpers

BF = torch.bfloat16
F32 = torch.float32
ACT = 1920          # saved activations per sample: outputs of mlp.0/2/4/6, feature_fn.0/2/4 (7 x 256) + rgb_fn.0 (128)
DZ = ACT + 16       # dz carries a 16-wide heads block [dsigma_pre, drgb_pre x3, 0 ...] after the 1920 layer columns
TILE = 128


def padded_rows(M):
    return (M + TILE - 1) // TILE * TILE


def untile(buf, M, features):
    """Tiled chunk-major training tensor (csrc/pack_layout.cuh) -> plain [M, features] row-major copy."""
    tiles = padded_rows(M) // TILE
    return buf.view(tiles, features // 8, TILE, 8).permute(0, 2, 1, 3).reshape(tiles * TILE, features)[:M]


def mlp_forward_train(model, o, d, ts):
    """Fused MLP in training form.  Returns sigma [N,S,1], rgb [N,S,3] and saved = (bf16 activations in the tiled
    chunk-major layout, ceil(N*S/128)*128 rows x 1920 features, flat; ReLU sign words for the dgrad kernel)."""
    N, S = ts.shape[0], ts.shape[1]
    sigma = torch.empty((N, S, 1), device=ts.device, dtype=F32)
    rgb = torch.empty((N, S, 3), device=ts.device, dtype=F32)
    acts = torch.empty((padded_rows(N * S) * ACT,), device=ts.device, dtype=BF)
    masks = torch.empty((padded_rows(N * S) * (ACT // 64),), device=ts.device, dtype=torch.int64)   # ReLU sign words for dgrad
    packed = model.packed_weights()
    with nat.timed_kernel("mlp_tc_kernel(train)", N * S):
        nat.check(nat.lib().nerf_mlp_forward_tc_train(nat.ptr(packed), nat.ptr(o), nat.ptr(d), nat.ptr(ts), N, S,
                                                      nat.ptr(sigma), nat.ptr(rgb), nat.ptr(acts), nat.ptr(masks), nat.stream()),
                  "nerf_mlp_forward_tc_train")
    return sigma, rgb, (acts, masks)


def composite_backward(sigma, rgb, ts, g_ray):
    N, S = sigma.shape[0], sigma.shape[1]
    dsig = torch.empty((N * S,), device=sigma.device, dtype=F32)
    drgb = torch.empty((N * S, 3), device=sigma.device, dtype=F32)
    g = nat.dev(g_ray, "g_ray")
    nat.check(nat.lib().nerf_composite_backward(nat.ptr(sigma), nat.ptr(rgb), nat.ptr(ts), nat.ptr(g), N, S,
                                                nat.ptr(dsig), nat.ptr(drgb), nat.stream()), "nerf_composite_backward")
    return dsig, drgb


FUSE_COMPOSITE_BACKWARD = True     # the compositing backward inside the dgrad kernel's producer warps (nerf_mlp_backward_tc_fused)
DETERMINISTIC_WGRAD = False        # True: one CTA per wgrad job instead of split-K + atomics - bit-reproducible gradients, ~16x slower
                                   # weight-gradient kernel; a parity-debugging mode


def mlp_backward(model, o, d, ts, sigma, rgb, acts, g_ray, accumulate_into_grad=False):
    """Gradients of one network's 20 parameters (state_dict order) given dL/d(ray colour) [N,3]: all hand-written kernels -
    tcgen05 dgrad chain with the compositing backward in its producer warps (mlp_tc_bwd3.cu), tcgen05 wgrad + bias sums
    (wgrad_tc.cu).  FUSE_COMPOSITE_BACKWARD = False keeps the compositing backward as its own launch (bit-identical dz)."""
    import ctypes
    acts, masks = acts
    N, S = ts.shape[0], ts.shape[1]
    M = N * S
    dz_t = torch.empty((padded_rows(M) * DZ,), device=ts.device, dtype=BF)
    if FUSE_COMPOSITE_BACKWARD and S <= 1024:
        g = nat.dev(g_ray, "g_ray")
        with nat.timed_kernel("mlp_tc_bwd_kernel", M):
            nat.check(nat.lib().nerf_mlp_backward_tc_fused(nat.ptr(model.packed_weights_t()), nat.ptr(masks), nat.ptr(sigma), nat.ptr(rgb),
                                                           nat.ptr(ts), nat.ptr(g), N, S, nat.ptr(dz_t), nat.stream()),
                      "nerf_mlp_backward_tc_fused")
    else:
        dsig, drgb = composite_backward(sigma, rgb, ts, g_ray)
        with nat.timed_kernel("mlp_tc_bwd_kernel", M):
            nat.check(nat.lib().nerf_mlp_backward_tc(nat.ptr(model.packed_weights_t()), nat.ptr(masks), nat.ptr(dsig), nat.ptr(drgb),
                                                     N, S, nat.ptr(dz_t), nat.stream()), "nerf_mlp_backward_tc")
    params = model.ordered_params()
    direct = accumulate_into_grad and all(p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == F32 for p in params)
    if direct:
        # wgrad ACCUMULATES with atomics: add straight into the existing .grad buffers (views of the optimiser's flat
        # gradient buffer) and hand autograd nothing to accumulate - 40 AccumulateGrad adds and a memset per step less
        grads = [p.grad for p in params]
    else:
        flat = torch.zeros(sum(p.numel() for p in params), device=ts.device, dtype=F32)
        grads, off = [], 0
        for p in params:
            grads.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    arr = (ctypes.c_void_p * 20)(*[g.data_ptr() for g in grads])
    with nat.timed_kernel("wgrad_tc_kernel", M):
        nat.check(nat.lib().nerf_wgrad_tc(nat.ptr(acts), nat.ptr(dz_t), nat.ptr(o), nat.ptr(d), nat.ptr(ts), N, S, arr,
                                          1 if DETERMINISTIC_WGRAD else 0, nat.stream()),
                  "nerf_wgrad_tc")
    return [None] * 20 if direct else grads


FUSE_STRATA = True          # K1 (stratified depths) inside the coarse network's kernel as well
FUSE_COMPOSITE = True       # network + compositing in one kernel where the sample counts allow it (64 / 128 / 192 / 256)


_stats_pool = {}            # (device, stream) -> [zeroed fp32 block, next free float]: one memset per 64 forwards instead of one each


def _zeroed_stats8(dv):
    """8 zeroed floats (the density statistics of both networks) cut from a pooled block.  Views handed out earlier keep their
    block alive, so the logged statistics of earlier forwards stay valid.  Under CUDA-graph capture the floats get their own
    memset node instead (a replayed graph would keep accumulating into a slice that is zeroed only once)."""
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros((8,), device=dv, dtype=F32)
    key = (dv, torch.cuda.current_stream(dv).cuda_stream)      # the memset is ordered on the stream that will use the floats
    ent = _stats_pool.get(key)
    if ent is None or ent[1] + 8 > ent[0].numel():
        ent = [torch.zeros((8 * 64,), device=dv, dtype=F32), 0]
        _stats_pool[key] = ent
    out = ent[0][ent[1]:ent[1] + 8]
    ent[1] += 8
    return out


RENDER_ONE_CALL = True      # inference passes go through nerf_render_forward (the three launches queued by ONE call of the C ABI)


def _render_one_call(net, o, d, rand, fine_out):
    """NeRFNetwork.forward for rendering as one call of the C ABI (csrc/render.cu): returns what forward_pass returns; the
    intermediates (coarse depths / weights, merged depths) are views of the call's workspace."""
    N, C, Fn = o.shape[0], net.coarse_samples, net.fine_samples
    dv = o.device
    u_c, eps, u_f = (nat.dev(t, "rand") for t in rand)
    lib = nat.lib()
    ws = torch.empty((lib.nerf_render_workspace_bytes(N, C, Fn),), device=dv, dtype=torch.uint8)
    al = lambda b: (b + 255) & ~255
    off, views = 0, []
    for shape in ((N, C, 1), (N, C, 1), (N, C + Fn, 1), (N,), (N,)):
        n = 4
        for k in shape:
            n *= k
        views.append(ws[off:off + n].view(torch.float32).view(shape))
        off += al(n)
    c_ts, c_w, ts, c_depth, c_acc = views
    c_rgb = torch.empty((N, 3), device=dv, dtype=F32)
    f_rgb = fine_out if fine_out is not None else torch.empty((N, 3), device=dv, dtype=F32)
    if f_rgb.shape != (N, 3) or f_rgb.dtype != F32 or not f_rgb.is_contiguous() or f_rgb.device != dv:
        raise ValueError("forward: `fine_out` must be a contiguous [N,3] fp32 tensor on the rays' device")
    depth = torch.empty((N,), device=dv, dtype=F32)
    acc = torch.empty((N,), device=dv, dtype=F32)
    stats8 = _zeroed_stats8(dv)
    t_base, step = nerf_helpers._strata(net.near, net.far, C, dv)
    q_base = nerf_helpers._queries(Fn, dv)
    if True:
        # near / far are NOT forwarded to the fine sampler upstream (nerf_model.py:114-115): its 2.0 / 6.0 defaults apply
        nat.check(lib.nerf_render_forward(nat.ptr(net.coarse_network.packed_weights()), nat.ptr(net.fine_network.packed_weights()),
                                          nat.ptr(o), nat.ptr(d), nat.ptr(u_c), nat.ptr(t_base), float(step), nat.ptr(eps), nat.ptr(u_f),
                                          nat.ptr(q_base), N, C, Fn, 2.0, 6.0, nat.ptr(c_rgb), nat.ptr(f_rgb), nat.ptr(depth), nat.ptr(acc),
                                          nat.ptr(stats8), nat.ptr(ws), nat.stream()), "nerf_render_forward")
        nat.launches += 2                      # three kernels of ours behind the one call
    c = {"weights": c_w, "rgb": c_rgb, "depth": c_depth, "acc": c_acc, "stats": stats8[:2], "norm": stats8[2], "ts": c_ts}
    f = {"weights": None, "rgb": f_rgb, "depth": depth, "acc": acc, "stats": stats8[4:6], "norm": stats8[6]}
    aux = {"c": c, "f": f, "c_ts": c_ts, "ts": ts, "c_sigma": None, "c_rgb": None, "f_sigma": None, "f_rgb": None,
           "c_acts": None, "f_acts": None}
    return c_rgb, f_rgb, aux


def forward_pass(net, o, d, rand, save, keep_samples=False, fine_out=None):
    """Shared by inference and training: returns (coarse_rgb, fine_rgb, aux dict).  With the fused kernel the per-sample
    sigma / rgb of an inference pass are only materialised when `keep_samples` (aux['c_sigma'] ... are None otherwise)."""
    N, C, Fn = o.shape[0], net.coarse_samples, net.fine_samples
    dv = o.device
    if rand is None:        # the reference's draw order and shapes (nerf_helpers.py:52,139,154)
        rand = (torch.rand((N, C), device=dv), torch.rand((N, 1), device=dv), torch.rand((N, Fn, 1), device=dv))
    u_c, eps, u_f = rand
    fused = FUSE_COMPOSITE and net.coarse_network.can_composite(C) and net.fine_network.can_composite(C + Fn)
    if (fused and FUSE_STRATA and RENDER_ONE_CALL and not save and not keep_samples and C + Fn <= 256
            and nat.kernel_events is None):        # (per-kernel event timing, bench.py's roofline, needs the launches one by one)
        return _render_one_call(net, o, d, rand, fine_out)
    if fused:       # network + compositing in one kernel; render keeps no per-sample outputs at all
        stats8 = _zeroed_stats8(dv)                              # density statistics of both networks
        if FUSE_STRATA:                                          # ... and the stratified depths are formed in that kernel too
            t_base, step = nerf_helpers._strata(net.near, net.far, C, dv)
            c = net.coarse_network.render_rays(o, d, None, want_weights=True, keep_samples=keep_samples, save=save, stats=stats8[:4],
                                               strata=(nat.dev(u_c, "u_c").reshape(N, C), t_base, step))
            c_ts = c["ts"]
        else:
            c_ts = net._coarse_ts(o, d, u_c)
            c = net.coarse_network.render_rays(o, d, c_ts, want_weights=True, keep_samples=keep_samples, save=save, stats=stats8[:4])
        c_sigma, c_rgb, c_acts = c["sigma"], c["rgb_samples"], c["saved"]
    else:
        c_ts = net._coarse_ts(o, d, u_c)
        if save:
            c_sigma, c_rgb, c_acts = mlp_forward_train(net.coarse_network, o, d, c_ts)
        else:
            c_sigma, c_rgb = net.coarse_network.forward_rays(o, d, c_ts)
            c_acts = None
        c = nerf_helpers.composite(c_sigma, c_rgb, c_ts)
    # near / far are NOT forwarded upstream (nerf_model.py:114-115): the sampler's 2.0 / 6.0 defaults apply
    if C + Fn <= 256:       # fine depths + merge sort in one launch
        ts = nerf_helpers.fine_depths_sorted(c["weights"], c_ts, Fn, rand=(eps, u_f))
    else:
        _, f_ts = nerf_helpers.inverse_transform_sampling(o, d, c["weights"], c_ts, Fn, rand=(eps, u_f))
        _, ts = nerf_helpers.merge_samples(o, d, f_ts, c_ts, want_points=False)
    if fused:
        f = net.fine_network.render_rays(o, d, ts, want_weights=False, keep_samples=keep_samples, save=save, stats=stats8[4:],
                                         out=fine_out)
        f_sigma, f_rgb, f_acts = f["sigma"], f["rgb_samples"], f["saved"]
    else:
        if save:
            f_sigma, f_rgb, f_acts = mlp_forward_train(net.fine_network, o, d, ts)
        else:
            f_sigma, f_rgb = net.fine_network.forward_rays(o, d, ts)
            f_acts = None
        f = nerf_helpers.composite(f_sigma, f_rgb, ts, want_weights=False)
    aux = {"c": c, "f": f, "c_ts": c_ts, "ts": ts, "c_sigma": c_sigma, "c_rgb": c_rgb, "f_sigma": f_sigma, "f_rgb": f_rgb,
           "c_acts": c_acts, "f_acts": f_acts}
    if fine_out is not None and f["rgb"] is not fine_out:      # two-launch path: the compositing kernel allocated its own output
        fine_out.copy_(f["rgb"])
    return c["rgb"], f["rgb"], aux


class NotDifferentiable(torch.autograd.Function):
    """Marks the outputs of a forward that has no hand-written backward (precision="fp32" or encoding sizes other than 10 / 4:
    the exact-fp32 CUDA-core kernel is forward-only).  The forward works as before; a `.backward()` through it fails HERE, with
    a message that names the cause, instead of deep inside torch with "element 0 of tensors does not require grad"."""

    @staticmethod
    def forward(ctx, why, anchor, *outputs):
        ctx.why = why
        return tuple(t.view_as(t) for t in outputs)

    @staticmethod
    def backward(ctx, *grads):
        raise RuntimeError(ctx.why)


class RenderFunction(torch.autograd.Function):
    """(o, d, u_c, eps, u_f, *40 parameters) -> (coarse_rgb_rays [N,3], fine_rgb_rays [N,3])."""

    @staticmethod
    def forward(ctx, net, o, d, u_c, eps, u_f, *params):
        rand = None if u_c is None else (u_c, eps, u_f)
        c_rgb, f_rgb, aux = forward_pass(net, o, d, rand, save=True)
        ctx.net, ctx.o, ctx.d, ctx.aux = net, o, d, aux
        net._publish(aux)
        return c_rgb, f_rgb

    @staticmethod
    def backward(ctx, g_c, g_f):
        net, o, d, a = ctx.net, ctx.o, ctx.d, ctx.aux
        gc = mlp_backward(net.coarse_network, o, d, a["c_ts"], a["c_sigma"], a["c_rgb"], a["c_acts"], g_c.contiguous(), True)
        if all(g is None for g in gc) and getattr(net, "on_coarse_grads_ready", None) is not None:
            net.on_coarse_grads_ready()      # data parallel: the coarse slice of the all-reduce overlaps the fine backward
        gf = mlp_backward(net.fine_network, o, d, a["ts"], a["f_sigma"], a["f_rgb"], a["f_acts"], g_f.contiguous(), True)
        ctx.aux = None
        return (None,) * 6 + tuple(gc) + tuple(gf)


class GraphedTrainStep:
    """One optimiser step of NeRFNetwork - batch draw, forward, the two MSE losses, backward, gradient sum over the data-parallel
    ranks, Adam, re-pack of the bf16 weight images - captured ONCE as a CUDA graph and replayed: ~35 kernel launches, the autograd
    engine and every eager op of `training_step` cost one `cudaGraphLaunch` per step (the kernels of a 4096-ray step take ~3.2 ms,
    the eager step 3.5 ms).  What changes from step to step lives in device memory: the image index (`img_idx`, filled by `step()`),
    torch's Philox offsets (torch.cuda.graph's capture-aware generator), the optimiser's step count and learning rate
    (optim.FlatAdam graph-safe form).

    images [n,H,W,3] uint8 / poses [n,4,4]: the device-resident training split (dataloader.SyntheticDataset.stacked()); `cropping`
    as dataloader.sample_random_coordinates.  The captured body is exactly Trainer.fit's eager body:
        grads.zero(); loss = model.training_step(batch, 0); loss.backward(); grads.all_reduce_mean(); optimizer.step()
    """

    def __init__(self, model, optimizer, grads, images, poses, focal, num_rays, cropping=False, track_grad_norm=False, warmup=3,
                 warmup_images=None):
        import dataloader
        self.model, self.optimizer, self.grads = model, optimizer, grads
        dev = images.device
        self.img_idx = torch.zeros((), device=dev, dtype=torch.int64)
        n_img, H, W, _ = images.shape
        self.n_img = n_img
        self.grad_norm = None

        def body():
            xs, ys = dataloader.sample_random_coordinates(num_rays, H, W, cropping=cropping, device=dev)
            o, d, rgb = dataloader.batch_rays(poses, self.img_idx, images, focal, xs, ys)
            grads.zero()
            loss = model.training_step({"origin": o[None], "direc": d[None], "rgb": rgb[None], "xs": xs[None], "ys": ys[None]}, 0)
            loss.backward()
            grads.all_reduce_mean()
            if track_grad_norm:
                self.grad_norm = grads.norm()
            optimizer.step()
            return loss.detach()

        self._body = body
        optimizer.graph_safe = True
        optimizer.sync_device_state()
        optimizer.set_device_step()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up on a side stream, as torch.cuda.graph asks for
            for k in range(warmup):
                if warmup_images:
                    self.img_idx.fill_(int(warmup_images[k % len(warmup_images)]))
                self.loss = body()
                if k == 0:
                    self.first_loss = self.loss.clone()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.warmup_steps = warmup                         # real optimiser steps: the caller counts them
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = body()
        optimizer._step -= 1                               # the captured call only recorded a step: nothing ran
        self._logged = dict(getattr(model, "logged", {}))

    def close(self):
        """Destroy the captured graph (and release its private memory pool).  REQUIRED before torch.distributed's process group is
        destroyed when the step contains NCCL work: tearing a communicator down while a live CUDA graph still references its
        kernels hangs (measured: tools/probe_nccl_graph_teardown.py).  The object holds a reference cycle (the captured body
        closes over it), so dropping the last reference is not enough."""
        if self.graph is not None:
            self.graph.reset()
        self.graph, self._body, self.loss = None, None, None

    def eager_step(self, image_index):
        """The same step run eagerly (launch by launch) - for per-kernel timing and for comparison with the replayed graph."""
        self.optimizer.sync_device_state()
        self.img_idx.fill_(int(image_index))
        return self._body()

    def step(self, image_index):
        """Replay with the batch drawn from image `image_index`; returns the (device, 0-d) loss of this step."""
        self.optimizer.sync_device_state()                 # the LR scheduler may have changed lr since the last step
        self.img_idx.fill_(int(image_index))
        self.graph.replay()
        self.optimizer._step += 1
        return self.loss
