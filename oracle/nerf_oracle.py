"""CPU oracle for the NeRF render-and-train hot path.  TEST INFRASTRUCTURE ONLY.

This is a restatement, in CPU float32 torch ops, of the arithmetic the reference
(NakuraMino/CSE-573-Minimal-NeRF, mounted read-only at /root/reference in the build container)
performs on the path SURVEY.md section 8 names.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product path never does
and fails loudly when the CUDA library is missing.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference modules in the
build container, feeds them seeded inputs with the `torch.rand` draws replaced by recorded numbers, and
commits the outputs as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below
against those files bit-for-bit (the MLP within 1e-6) and against the known-answer values held by the
reference's own unit tests (tests/nerf_helpers_test.py:16-63, tests/nerf_model_test.py:41-63,
tests/dataloader_test.py:39-41).

Differences from the reference that are deliberate:
  * random numbers are arguments (`u_*`, `eps`), never drawn here; the reference calls `torch.rand`
    three times per forward with shapes [N,C], [N,1], [N,F,1] (nerf_helpers.py:52,139,154);
  * everything runs on CPU tensors (the reference allocates on a module-global device,
    nerf_helpers.py:16) - the CPU arithmetic (sequential cumsum, true division) is the contract;
  * depth and accumulated opacity, which the reference never outputs, are derived from its weights.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- camera / rays


def pose_spherical(theta, phi, radius):
    """c2w for the 360 orbit; follows nerf_helpers.py:258-284 (float64 numpy trig -> float32 4x4s)."""
    def m(rows):
        return torch.tensor(rows, dtype=torch.float32)
    ph, th = phi / 180.0 * np.pi, theta / 180.0 * np.pi
    c2w = m([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]])
    c2w = m([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]]) @ c2w
    c2w = m([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]]) @ c2w
    return torch.tensor([[-1.0, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w


def focal_from_fov(width, cam_angle_x):
    """nerf_helpers.py:180 / dataloader.py:121."""
    return 0.5 * width / np.tan(0.5 * cam_angle_x)


def get_rays(H, W, focal, c2w, xs=None, ys=None):
    """Pinhole rays, dataloader.py:36-43.  With xs/ys ([n] int64, column / row) only those pixels are
    produced (the gather of dataloader.py:150-152).  Returns (o[...,3], d[...,3])."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    if xs is None:
        col = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)
        row = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
    else:
        col, row = xs.to(torch.float32), ys.to(torch.float32)
    # focal arrives as a float64 scalar; tensor/scalar division on CPU divides by fl32(focal)
    d0 = (col - W * .5) / focal
    d1 = -(row - H * .5) / focal
    d2 = -torch.ones_like(d0)
    R = c2w[:3, :3]
    d = torch.stack([(d0 * R[k, 0] + d1 * R[k, 1]) + d2 * R[k, 2] for k in range(3)], -1)
    o = c2w[:3, 3].expand(d.shape)
    return o, d

# ----------------------------------------------------------------------------- sampling


def strata_origins(near, far, C):
    """The `torch.arange(near, far, step)` of nerf_helpers.py:50-51 (a torch library call, float32)."""
    return torch.arange(near, far, (far - near) / C)


def query_grid(F_):
    """The `torch.arange(0, 1, 1/F)` of nerf_helpers.py:140."""
    return torch.arange(0, 1, 1 / F_)


def coarse_samples(o, d, u, near=2.0, far=6.0):
    """Stratified depths and points, nerf_helpers.py:28-56.  u: [N,C] uniforms.
    Returns samples [N,C,3], ts [N,C,1]."""
    N, C = u.shape
    step = (far - near) / C
    ts = strata_origins(near, far, C)[None, :] + u * step          # scalar is rounded to fp32 first
    ts = ts[..., None]
    return d[:, None, :] * ts + o[:, None, :], ts


def deltas(ts):
    """Adjacent differences with a 1e10 tail, nerf_helpers.py:58-73."""
    last = torch.full_like(ts[:, :1, :], 1e10, dtype=torch.float32)
    return torch.cat([ts[:, 1:, :] - ts[:, :-1, :], last], dim=1)


def weights(sigma, delta):
    """w_i = exp(sum_{j<i} -sigma_j delta_j) * (1 - exp(-sigma_i delta_i)), nerf_helpers.py:75-91."""
    nds = -1 * sigma * delta
    shifted = torch.cat([torch.zeros_like(nds[:, :1, :]), nds[:, :-1, :]], dim=1)
    trans = torch.exp(torch.cumsum(shifted, dim=1))
    return (1 - torch.exp(nds)) * trans


def ray_color(w, rgb):
    """sum_i w_i c_i, nerf_helpers.py:93-104."""
    return torch.sum(w * rgb, dim=1)


def depth_and_acc(w, ts):
    """Derived (not in the reference): expected depth sum w t and accumulated opacity sum w."""
    return torch.sum(w * ts, dim=1), torch.sum(w, dim=1)


def fine_samples(o, d, w, ts, eps, u, near=2.0, far=6.0, return_idx=False):
    """Inverse-CDF sampling, nerf_helpers.py:106-156.  w, ts: [N,C,1]; eps: [N,1]; u: [N,F,1].
    near/far default to 2/6 because the reference's caller never forwards them (nerf_model.py:114-115)."""
    N, C, _ = ts.shape
    F_ = u.shape[1]
    cdf = torch.cumsum(w, dim=1)
    cdf = (cdf / cdf[:, -1, None]).squeeze(-1)
    q = query_grid(F_)[None, :].expand(N, F_) + eps / F_
    lo_idx = torch.searchsorted(cdf, q.contiguous())[..., None]
    bounds = torch.cat([torch.full((N, 1, 1), near), ts, torch.full((N, 1, 1), far)], dim=1)
    lo = torch.gather(bounds, 1, lo_idx)
    hi = torch.gather(bounds, 1, lo_idx + 1)
    fts = lo + (hi - lo) * u
    pts = o[:, None, :] + fts * d[:, None, :]
    if return_idx:
        return pts, fts, lo_idx.squeeze(-1)
    return pts, fts


def merge_sorted(pts_a, ts_a, pts_b, ts_b):
    """Concatenate (fine first, then coarse), sort by depth, gather points: nerf_model.py:116-120."""
    pts = torch.cat([pts_a, pts_b], dim=1)
    ts = torch.cat([ts_a, ts_b], dim=1)
    ts, order = torch.sort(ts, dim=1)
    return torch.gather(pts, 1, order.expand(pts.shape)), ts

# ----------------------------------------------------------------------------- network


def positional_encoding(x, L):
    """Per frequency i: [cos(2^i pi x) (C values), sin(2^i pi x) (C values)], nerf_model.py:19-33."""
    out = []
    for i in range(L):
        out += [torch.cos(2 ** i * torch.pi * x), torch.sin(2 ** i * torch.pi * x)]
    return torch.cat(out, dim=-1)


def mlp_forward(sd, prefix, samples, direc, position_dim=10, direction_dim=4):
    """One NeRFModel forward (nerf_model.py:362-389) from a state_dict slice.
    samples [N,S,3], direc [N,3] -> sigma [N,S,1], rgb [N,S,3]."""
    def lin(x, name):
        return F.linear(x, sd[f"{prefix}.{name}.weight"], sd[f"{prefix}.{name}.bias"])
    unit = direc / torch.linalg.norm(direc, dim=1, keepdim=True)
    unit = unit[:, None, :].expand(samples.shape)
    pe_x = positional_encoding(samples / math.pi, position_dim)
    pe_d = positional_encoding(unit, direction_dim)
    h = pe_x
    for name in ("mlp.0", "mlp.2", "mlp.4", "mlp.6"):
        h = torch.relu(lin(h, name))
    h = torch.cat([h, pe_x], dim=-1)
    h = torch.relu(lin(h, "feature_fn.0"))
    h = torch.relu(lin(h, "feature_fn.2"))
    feat = lin(h, "feature_fn.4")
    sigma = torch.relu(lin(feat, "density_fn.0"))
    r = torch.relu(lin(torch.cat([feat, pe_d], dim=-1), "rgb_fn.0"))
    rgb = torch.sigmoid(lin(r, "rgb_fn.2"))
    return sigma, rgb


def network_forward(sd, o, d, u_c, eps, u_f, near=2.0, far=6.0, position_dim=10, direction_dim=4,
                    full=False):
    """NeRFNetwork.forward (nerf_model.py:89-132) with the three random draws supplied.
    u_c [N,C], eps [N,1], u_f [N,F,1].  Returns the reference's dict; with full=True also every
    intermediate the parity tests look at."""
    c_pts, c_ts = coarse_samples(o, d, u_c, near, far)
    c_sigma, c_rgb = mlp_forward(sd, "coarse_network", c_pts, d, position_dim, direction_dim)
    c_w = weights(c_sigma, deltas(c_ts))
    c_ray = ray_color(c_w, c_rgb)
    f_pts, f_ts = fine_samples(o, d, c_w, c_ts, eps, u_f)            # 2.0 / 6.0 defaults, as upstream
    pts, ts = merge_sorted(f_pts, f_ts, c_pts, c_ts)
    f_sigma, f_rgb = mlp_forward(sd, "fine_network", pts, d, position_dim, direction_dim)
    f_w = weights(f_sigma, deltas(ts))
    f_ray = ray_color(f_w, f_rgb)
    out = {"fine_rgb_rays": f_ray, "coarse_rgb_rays": c_ray}
    if full:
        depth, acc = depth_and_acc(f_w, ts)
        out.update(coarse_ts=c_ts, coarse_sigma=c_sigma, coarse_rgb=c_rgb, coarse_weights=c_w,
                   fine_ts=f_ts, ts=ts, fine_sigma=f_sigma, fine_rgb=f_rgb, fine_weights=f_w,
                   depth=depth, acc=acc,
                   stats=torch.stack([torch.linalg.norm(c_sigma), (c_sigma != 0).sum().float(),
                                      torch.linalg.norm(f_sigma), (f_sigma != 0).sum().float()]))
    return out


def training_loss(sd, o, d, target, u_c, eps, u_f, **kw):
    """coarse MSE + fine MSE (nerf_model.py:159-161)."""
    out = network_forward(sd, o, d, u_c, eps, u_f, **kw)
    return F.mse_loss(out["coarse_rgb_rays"], target) + F.mse_loss(out["fine_rgb_rays"], target), out


def loss_and_grads(sd, o, d, target, u_c, eps, u_f, **kw):
    """Loss and d loss / d parameter for all 40 tensors through CPU autograd (what PL's backward does)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss, out = training_loss(leaf, o, d, target, u_c, eps, u_f, **kw)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items()}, {k: v.detach() for k, v in out.items()}

# ----------------------------------------------------------------------------- view synthesis


def view_reconstruction(sd, all_o, all_d, rand_fn, N=4096, C=64, F_=128, near=2.0, far=6.0):
    """Chunked full-image render to uint8, nerf_helpers.py:189-210.  rand_fn(chunk_index, n) returns the
    (u_c [n,C], eps [n,1], u_f [n,F,1]) triple the reference would have drawn for that chunk."""
    H, W, _ = all_o.shape
    o, d = all_o.reshape(H * W, 3), all_d.reshape(H * W, 3)
    rows = []
    with torch.no_grad():
        for ci, i in enumerate(range(0, H * W, N)):
            oo, dd = o[i:i + N], d[i:i + N]
            u_c, eps, u_f = rand_fn(ci, oo.shape[0])
            rows.append(network_forward(sd, oo, dd, u_c, eps, u_f, near, far)["fine_rgb_rays"].numpy())
    im = np.concatenate(rows, axis=0).reshape(H, W, 3)
    im *= 255
    return np.clip(im, 0, 255).astype(np.uint8)


def psnr_uint8(a, b):
    """10 log10(255^2 / MSE) over uint8 images (skimage's peak_signal_noise_ratio, score.py:36)."""
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
