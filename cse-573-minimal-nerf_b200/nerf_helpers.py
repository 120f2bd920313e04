"""B200 drop-in for the reference's `nerf_helpers.py` (same names, arguments, defaults and return shapes).

Every function body is a launch of a hand-written sm_100a kernel from libnerf_b200.so; the random draws
stay `torch.rand` calls in the reference's order and shapes so a seeded run consumes the generator the
same way.  Inputs must be CUDA tensors - there is no CPU path.

Reference counterparts (file:line in NakuraMino/CSE-573-Minimal-NeRF):
  fix_batchify 18-26, generate_coarse_samples 28-56, generate_deltas 58-73,
  calculate_unnormalized_weights 75-91, estimate_ray_color 93-104, inverse_transform_sampling 106-156,
  generate_360_view_synthesis 162-187, view_reconstruction 189-210, pose_spherical 258-284.
"""
from pathlib import Path

import numpy as np
import torch

import _native as nat
import dataloader

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

_grid_cache = {}


def _strata(near, far, C, dev):
    """`torch.arange(near, far, step)` (nerf_helpers.py:50-51) evaluated on the CPU in fp32 - the way the reference evaluates it
    when it runs on the CPU, which is where the parity vectors were recorded - and cached on the device.  Upstream passes `device=device`, i.e. a CUDA arange when a GPU is
    present; the two can differ by 1 ulp when `step` is not exactly representable (the default 2..6 / 64 gives step = 2^-4: exact)."""
    key = ("t", float(near), float(far), int(C), str(dev))
    if key not in _grid_cache:
        step = (far - near) / C
        base = torch.arange(near, far, step)
        if base.shape[0] != C:
            raise RuntimeError(f"arange({near}, {far}, {step}) has {base.shape[0]} elements, expected {C} "
                               "(the reference's broadcast_to would fail the same way)")
        _grid_cache[key] = (base.to(dev), float(np.float32(step)))
    return _grid_cache[key]


def _queries(F, dev):
    key = ("q", int(F), str(dev))
    if key not in _grid_cache:
        _grid_cache[key] = torch.arange(0, 1, 1 / F)[:F].contiguous().to(dev)
    return _grid_cache[key]


def fix_batchify(batch):
    """Drop the DataLoader's leading 1 from every tensor of the batch, in place."""
    for key in batch:
        batch[key] = batch[key].squeeze(0)


@nat.device_guard
def generate_coarse_samples(o_rays, d_rays, num_samples, near=2.0, far=6.0, rand=None):
    """Stratified samples: returns (samples [N,num_samples,3], ts [N,num_samples,1]).
    `rand` ([N,num_samples] uniforms) replaces the internal torch.rand draw when given."""
    o, d = nat.dev(o_rays, "o_rays"), nat.dev(d_rays, "d_rays")
    N = o.shape[0]
    t_base, step = _strata(near, far, num_samples, o.device)
    u = torch.rand((N, num_samples), device=o.device) if rand is None else nat.dev(rand, "rand").reshape(N, num_samples)
    samples = torch.empty((N, num_samples, 3), device=o.device, dtype=torch.float32)
    ts = torch.empty((N, num_samples, 1), device=o.device, dtype=torch.float32)
    nat.check(nat.lib().nerf_coarse_sample(nat.ptr(o), nat.ptr(d), nat.ptr(u), nat.ptr(t_base), step, N, num_samples,
                                           nat.ptr(samples), nat.ptr(ts), nat.stream()), "nerf_coarse_sample")
    return samples, ts


@nat.device_guard
def generate_deltas(ts):
    """delta_i = t_{i+1} - t_i with a 1e10 tail.  ts: [N,S,1] (any numeric dtype, promoted to fp32)."""
    t = nat.dev(ts, "ts")
    N, S, _ = t.shape
    out = torch.empty((N, S, 1), device=t.device, dtype=torch.float32)
    nat.check(nat.lib().nerf_deltas(nat.ptr(t), N, S, nat.ptr(out), nat.stream()), "nerf_deltas")
    return out


@nat.device_guard
def calculate_unnormalized_weights(density, deltas):
    """w = T (1 - exp(-density * delta)); density, deltas: [N,S,1]."""
    sg, dl = nat.dev(density, "density"), nat.dev(deltas, "deltas")
    N, S, _ = sg.shape
    out = torch.empty((N, S, 1), device=sg.device, dtype=torch.float32)
    nat.check(nat.lib().nerf_weights(nat.ptr(sg), nat.ptr(dl), N, S, nat.ptr(out), nat.stream()), "nerf_weights")
    return out


@nat.device_guard
def estimate_ray_color(weights, rgb):
    """sum_i w_i rgb_i.  weights [N,S,1], rgb [N,S,3] -> [N,3]."""
    w, c = nat.dev(weights, "weights"), nat.dev(rgb, "rgb")
    N, S, _ = w.shape
    out = torch.empty((N, 3), device=w.device, dtype=torch.float32)
    nat.check(nat.lib().nerf_ray_color(nat.ptr(w), nat.ptr(c), N, S, nat.ptr(out), nat.stream()), "nerf_ray_color")
    return out


@nat.device_guard
def composite(density, rgb, ts, want_weights=True):
    """deltas + weights + ray colour + depth + opacity in one launch (what NeRFNetwork.forward needs).
    Returns dict(weights [N,S,1], rgb [N,3], depth [N], acc [N], stats [2] = (sum sigma^2, count sigma != 0),
    norm = sqrt(sum sigma^2), a 0-d view written by the same launch)."""
    sg, c, t = nat.dev(density, "density"), nat.dev(rgb, "rgb"), nat.dev(ts, "ts")
    N, S = sg.shape[0], sg.shape[1]
    dv = sg.device
    w = torch.empty((N, S, 1), device=dv, dtype=torch.float32) if want_weights else None
    col = torch.empty((N, 3), device=dv, dtype=torch.float32)
    depth = torch.empty((N,), device=dv, dtype=torch.float32)
    acc = torch.empty((N,), device=dv, dtype=torch.float32)
    stats = torch.zeros((4,), device=dv, dtype=torch.float32)
    nat.check(nat.lib().nerf_composite(nat.ptr(sg), nat.ptr(c), nat.ptr(t), N, S, None, nat.ptr(w), nat.ptr(col),
                                       nat.ptr(depth), nat.ptr(acc), nat.ptr(stats), nat.stream()), "nerf_composite")
    return {"weights": w, "rgb": col, "depth": depth, "acc": acc, "stats": stats[:2], "norm": stats[2]}


@nat.device_guard
def inverse_transform_sampling(o_rays, d_rays, weights, ts, num_samples, near=2.0, far=6.0, rand=None,
                               return_idx=False):
    """Inverse-CDF fine samples: returns (fine_samples [N,num_samples,3], fine_ts [N,num_samples,1]).
    `rand` = (eps [N,1], u [N,num_samples,1]) replaces the two internal torch.rand draws when given."""
    o, d = nat.dev(o_rays, "o_rays"), nat.dev(d_rays, "d_rays")
    w, t = nat.dev(weights, "weights"), nat.dev(ts, "ts")
    N, C, _ = t.shape
    if rand is None:
        eps = torch.rand((N, 1), device=o.device)
        u = torch.rand((N, num_samples, 1), device=o.device)
    else:
        eps, u = nat.dev(rand[0], "rand[0]"), nat.dev(rand[1], "rand[1]")
    q = _queries(num_samples, o.device)
    pts = torch.empty((N, num_samples, 3), device=o.device, dtype=torch.float32)
    fts = torch.empty((N, num_samples, 1), device=o.device, dtype=torch.float32)
    idx = torch.empty((N, num_samples), device=o.device, dtype=torch.int64) if return_idx else None
    nat.check(nat.lib().nerf_fine_sample(nat.ptr(o), nat.ptr(d), nat.ptr(w), nat.ptr(t), nat.ptr(eps), nat.ptr(u), nat.ptr(q),
                                         N, C, num_samples, float(near), float(far), nat.ptr(pts), nat.ptr(fts), nat.ptr(idx),
                                         nat.stream()), "nerf_fine_sample")
    return (pts, fts, idx) if return_idx else (pts, fts)


@nat.device_guard
def fine_depths_sorted(weights, ts, num_samples, near=2.0, far=6.0, rand=None):
    """inverse_transform_sampling + merge_samples in one launch, depths only (what NeRFNetwork.forward needs between the two
    networks, nerf_model.py:114-120): returns the sorted [N, C+num_samples, 1] depths, bit-identical to the two calls.
    C + num_samples <= 256."""
    w, t = nat.dev(weights, "weights"), nat.dev(ts, "ts")
    N, C, _ = t.shape
    if rand is None:
        eps = torch.rand((N, 1), device=t.device)
        u = torch.rand((N, num_samples, 1), device=t.device)
    else:
        eps, u = nat.dev(rand[0], "rand[0]"), nat.dev(rand[1], "rand[1]")
    q = _queries(num_samples, t.device)
    out = torch.empty((N, C + num_samples, 1), device=t.device, dtype=torch.float32)
    nat.check(nat.lib().nerf_fine_sample_merge(nat.ptr(w), nat.ptr(t), nat.ptr(eps), nat.ptr(u), nat.ptr(q), N, C, num_samples,
                                               float(near), float(far), nat.ptr(out), nat.stream()), "nerf_fine_sample_merge")
    return out


@nat.device_guard
def merge_samples(o_rays, d_rays, fine_ts, coarse_ts, want_points=True):
    """The concat + sort + gather of NeRFNetwork.forward (nerf_model.py:116-120) as one kernel.
    Returns (samples [N,A+B,3] or None, ts [N,A+B,1])."""
    o, d = nat.dev(o_rays, "o_rays"), nat.dev(d_rays, "d_rays")
    ta, tb = nat.dev(fine_ts, "fine_ts"), nat.dev(coarse_ts, "coarse_ts")
    N, A, B = ta.shape[0], ta.shape[1], tb.shape[1]
    ts = torch.empty((N, A + B, 1), device=o.device, dtype=torch.float32)
    pts = torch.empty((N, A + B, 3), device=o.device, dtype=torch.float32) if want_points else None
    nat.check(nat.lib().nerf_merge_sort(nat.ptr(o), nat.ptr(d), nat.ptr(ta), A, nat.ptr(tb), B, N, nat.ptr(ts), nat.ptr(pts),
                                        nat.stream()), "nerf_merge_sort")
    return pts, ts


# ------------------------------------------------------------------------------ view synthesis

def generate_360_view_synthesis(model, save_dir: Path, epoch, height=800, width=800,
                                radius=4.0, cam_angle_x=0.6911112070083618, N=4096,
                                num_poses=40):
    """Render `num_poses` views on a 360 degree orbit and save them as SAVE_DIR/EPOCH-360.gif."""
    save_dir = Path(save_dir)
    assert save_dir.exists() and save_dir.is_dir()
    focal = 0.5 * width / np.tan(0.5 * cam_angle_x)
    views = []
    for angle in np.linspace(-180, 180, num_poses + 1)[:-1]:
        o_rays, d_rays = dataloader.get_rays(height, width, focal, pose_spherical(angle, -30, radius), device=device)
        views.append(view_reconstruction(model, o_rays, d_rays, N=N))
    import multi_gpu
    if multi_gpu.world()[0] == 0:           # every rank holds all frames (all-gathered slabs); one of them writes the file
        dataloader.write_gif(Path(save_dir, f'{epoch}-360.gif'), views)
    return views


# The reference pushes N rays at a time through the network because eager PyTorch holds ~2.5 MB of activations per ray
# (nerf_helpers.py:204-206).  The fused kernels keep a sample's activations on the SM and need ~2.2 KB of HBM per ray (uniforms,
# depths, coarse weights), and their persistent CTAs loop over any number of ray groups, so consecutive chunks are GROUPED: up to
# RAYS_PER_LAUNCH rays go through one set of launches (3 torch.rand draws + coarse kernel + sampler + fine kernel = 6 launches
# for a whole 800x800 frame instead of 157 x 6, and one pipeline fill / drain per kernel instead of 157).  N keeps its meaning
# as the upper bound when it is larger than RAYS_PER_LAUNCH; `RAYS_PER_LAUNCH = None` restores one launch set per N-ray chunk.
# The uniforms are the same `torch.rand` calls in the same order, just with more rows per call.
RAYS_PER_LAUNCH = 1 << 20


def render_rays_chunked(model, o, d, N=4096, out=None):
    """The chunk loop of view_reconstruction (nerf_helpers.py:196-206): `model.forward` over all rays, fine colours gathered
    into one [n,3] tensor (`out`, written in place by the fine network's kernel)."""
    n = o.shape[0]
    if out is None:
        out = torch.empty((n, 3), device=o.device, dtype=torch.float32)
    direct = hasattr(model, "coarse_network") and hasattr(model, "fine_network") and out.is_contiguous()
    step = N if (RAYS_PER_LAUNCH is None or not direct) else max(N, RAYS_PER_LAUNCH)
    with torch.no_grad():
        for i in range(0, n, step):
            if direct:
                model.forward(o[i:i + step], d[i:i + step], fine_out=out[i:i + step])      # the kernel writes into the frame buffer
            else:
                out[i:i + step] = model.forward(o[i:i + step], d[i:i + step])['fine_rgb_rays']
    return out


_pinned = {}


def _to_host_u8(image):
    """uint8 [H,W,3] device image -> numpy array through a cached PINNED staging buffer (an asynchronous copy + one stream
    synchronisation instead of a pageable-memory `.cpu()`); the returned array owns its memory."""
    key = (tuple(image.shape), image.device.index)
    if key not in _pinned:
        _pinned[key] = torch.empty(image.shape, dtype=torch.uint8).pin_memory()
    stage = _pinned[key]
    stage.copy_(image, non_blocking=True)
    torch.cuda.current_stream(image.device).synchronize()
    return stage.numpy().copy()


def view_reconstruction(model, all_o_rays, all_d_rays, N=4096):
    """Render every ray of an [H,W,3] ray grid through `model`; returns uint8 [H,W,3] (host, numpy).
    Under torch.distributed (one process per GPU) every rank renders a contiguous slab of the rays and the uint8
    slabs are all-gathered, so each rank returns the whole image."""
    import multi_gpu
    o_all, d_all = nat.dev(all_o_rays, "all_o_rays"), nat.dev(all_d_rays, "all_d_rays")
    with torch.cuda.device(o_all.device):
        # (x * 255).clamp(0, 255).to(uint8) truncates like numpy's astype(uint8) upstream (nerf_helpers.py:207-210)
        return _to_host_u8(multi_gpu.sharded_render(lambda o, d: render_rays_chunked(model, o, d, N), o_all, d_all))


def torch_to_numpy(torch_tensor, is_normalized_image=False):
    """(...CHW) tensor -> (...HWC) numpy array; rescaled to [0,255] if it was a normalised image."""
    arr = torch_tensor.detach().cpu().clone().numpy()
    if arr.ndim >= 4:
        arr = np.moveaxis(arr, [-3, -2, -1], [-1, -3, -2])
    if is_normalized_image:
        arr = np.clip(arr * 255, 0, 255)
    return arr


# ------------------------------------------------------------------------------ orbit poses (host side)

def trans_t(t):
    return torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=torch.float32)


def rot_phi(phi):
    c, s = np.cos(phi), np.sin(phi)
    return torch.tensor([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=torch.float32)


def rot_theta(th):
    c, s = np.cos(th), np.sin(th)
    return torch.tensor([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=torch.float32)


def pose_spherical(theta, phi, radius):
    """Camera-to-world matrix on a sphere of `radius` (degrees), the orbit bmild/nerf renders."""
    flip = torch.tensor([[-1.0, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return flip @ (rot_theta(theta / 180. * np.pi) @ (rot_phi(phi / 180. * np.pi) @ trans_t(radius)))
