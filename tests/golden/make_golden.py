"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    cd /root/repo && python tests/golden/make_golden.py

Imports /root/reference/{nerf_helpers,nerf_model,dataloader}.py under two stub packages
(tests/golden/stubs: pytorch_lightning and imageio are not installed and there is no network), pins
`nerf_helpers.device` to CPU, and replaces `torch.rand` by a recorded stream
(`synthetic.uniforms(seed0 + call_index, shape)`) so the same numbers can be injected into the oracle
and the CUDA path.  /root/reference does not exist on the GPU box; only the .npz files travel.
"""
import hashlib
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(HERE / "stubs"))
sys.path.insert(0, "/root/reference")
sys.path.append(str(ROOT / "cse-573-minimal-nerf_b200"))      # `synthetic` only: the reference's same-named modules must win
os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np
import torch

import synthetic                      # ours (deterministic inputs)
import dataloader as ref_data         # reference
import nerf_helpers as ref_helpers    # reference
import nerf_model as ref_model        # reference

assert ref_helpers.__file__.startswith("/root/reference"), ref_helpers.__file__
ref_helpers.device = torch.device("cpu")
torch.set_num_threads(os.cpu_count())

_real_rand = torch.rand


class RandStream:
    """torch.rand replacement: call k returns synthetic.uniforms(seed0 + k, shape)."""

    def __init__(self, seed0):
        self.seed0, self.k = seed0, 0

    def __call__(self, *shape, **kw):
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        out = torch.from_numpy(synthetic.uniforms(self.seed0 + self.k, tuple(shape)))
        self.k += 1
        return out

    def __enter__(self):
        torch.rand = self
        return self

    def __exit__(self, *a):
        torch.rand = _real_rand


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        out[k] = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"wrote {name}.npz: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in out.items()))


def load_model(sd):
    net = ref_model.NeRFNetwork()
    net.load_state_dict(sd)
    return net


def orbit_rays(theta, n, seed, H=800, W=800):
    """n random pixels of an 800x800 orbit camera through the reference's own get_rays."""
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    c2w = ref_helpers.pose_spherical(theta, -30.0, 4.0)
    o, d = ref_data.get_rays(H, W, focal, c2w)
    xs = T((synthetic.uniforms(seed, (n,)) * W).astype(np.int64))
    ys = T((synthetic.uniforms(seed + 1, (n,)) * H).astype(np.int64))
    return o[ys, xs, :].contiguous(), d[ys, xs, :].contiguous(), xs, ys, c2w, focal


def gen_rays():
    arrs = {}
    focal800 = 0.5 * 800 / np.tan(0.5 * 0.6911112070083618)
    poses = [torch.Tensor([[1, 0, 0, .5], [0, 1, 0, .5], [0, 0, 1, .5], [0, 0, 0, 1]])]
    angles = [-180.0, -63.0, 90.0]
    poses += [ref_helpers.pose_spherical(a, -30.0, 4.0) for a in angles]
    arrs["poses"] = torch.stack(poses)
    arrs["angles"] = np.array(angles)
    arrs["focal800"] = np.float64(focal800)
    sums = []
    for k, p in enumerate(poses):
        o, d = ref_data.get_rays(48, 64, 77.25, p)              # H != W catches x/y mix-ups
        arrs[f"d_small_{k}"] = d
        arrs[f"o_small_{k}"] = o.contiguous()
        o, d = ref_data.get_rays(800, 800, focal800, p)
        sums.append(hashlib.sha256((d.contiguous().numpy() + 0.0).tobytes()).hexdigest())   # +0.0: canonical zero sign
        arrs[f"d800_row0_{k}"] = d[0]
        arrs[f"d800_row799_{k}"] = d[799]
    arrs["d800_sha256"] = np.array(sums)
    arrs["focal_fixture"] = np.float64(0.5 * 800 / np.tan(0.5 * 0.6))   # tests/dataloader_test.py:39-41
    save("rays", **arrs)


def gen_coarse():
    arrs = {}
    o, d, *_ = orbit_rays(20.0, 256, 11)
    with RandStream(100):
        pts, ts = ref_helpers.generate_coarse_samples(o, d, 64, 2.0, 6.0)
    arrs.update(o=o, d=d, pts=pts, ts=ts)
    o2, d2 = o[:5].contiguous(), d[:5].contiguous()
    with RandStream(110):
        pts2, ts2 = ref_helpers.generate_coarse_samples(o2, d2, 7, 0.3, 5.1)   # step not a power of two
    arrs.update(pts_odd=pts2, ts_odd=ts2)
    save("coarse", **arrs)


def sigma_patterns(seed, N, S):
    """relu(N(-2,3)) densities; every 5th ray all-zero; every 7th ray opaque early."""
    s = np.maximum(synthetic.normals(seed, (N, S, 1)) * 3.0 - 2.0, 0.0).astype(np.float32)
    s[::5] = 0.0
    s[3::7, S // 4] = 80.0
    return T(s)


def gen_composite():
    arrs = {}
    for S in (64, 192):
        N = 64
        step = 4.0 / S
        ts = T((2.0 + step * (np.arange(S)[None, :] + synthetic.uniforms(200 + S, (N, S)))).astype(np.float32))[..., None]
        sigma = sigma_patterns(210 + S, N, S)
        rgb = T(synthetic.uniforms(220 + S, (N, S, 3)))
        dl = ref_helpers.generate_deltas(ts)
        w = ref_helpers.calculate_unnormalized_weights(sigma, dl)
        col = ref_helpers.estimate_ray_color(w, rgb)
        arrs.update({f"ts_{S}": ts, f"sigma_{S}": sigma, f"rgb_{S}": rgb, f"deltas_{S}": dl,
                     f"weights_{S}": w, f"color_{S}": col})
    # the reference's own known-answer test inputs (tests/nerf_helpers_test.py:16-21)
    dl = torch.full((1, 5, 1), 0.2)
    sg = torch.Tensor([0, 50, 1, 0.3, 1]).view(dl.shape)
    arrs["kat_weights"] = ref_helpers.calculate_unnormalized_weights(sg, dl)
    save("composite", **arrs)


def gen_fine():
    arrs = {}
    N, C, Fn = 256, 64, 128
    o, d, *_ = orbit_rays(-100.0, N, 31)
    with RandStream(300):
        c_pts, c_ts = ref_helpers.generate_coarse_samples(o, d, C, 2.0, 6.0)
    sigma = sigma_patterns(310, N, C)
    w = ref_helpers.calculate_unnormalized_weights(sigma, ref_helpers.generate_deltas(c_ts))
    w[7] = 0.0
    w[7, 40] = 1.0e-30                                    # denormal-scale cdf
    recorded = {}
    real_ss = torch.searchsorted

    def spy(*a, **k):
        recorded["idx"] = real_ss(*a, **k)
        return recorded["idx"]
    torch.searchsorted = spy
    try:
        with RandStream(320):
            f_pts, f_ts = ref_helpers.inverse_transform_sampling(o, d, w, c_ts, Fn)
    finally:
        torch.searchsorted = real_ss
    # the merge of nerf_model.py:116-120 is inline in NeRFNetwork.forward; it is captured in gen_network.
    arrs.update(o=o, d=d, c_ts=c_ts, c_pts=c_pts, w=w, f_pts=f_pts, f_ts=f_ts, idx=recorded["idx"])
    # small odd shape: C=7, F=5
    with RandStream(330):
        c_pts2, c_ts2 = ref_helpers.generate_coarse_samples(o[:9].contiguous(), d[:9].contiguous(), 7, 2.0, 6.0)
    w2 = T(synthetic.uniforms(331, (9, 7, 1)))
    w2[2] = 0.0
    with RandStream(332):
        f_pts2, f_ts2 = ref_helpers.inverse_transform_sampling(o[:9].contiguous(), d[:9].contiguous(), w2, c_ts2, 5)
    arrs.update(c_ts_odd=c_ts2, w_odd=w2, f_pts_odd=f_pts2, f_ts_odd=f_ts2)
    save("fine", **arrs)


def gen_pe_mlp():
    arrs = {}
    x = T(synthetic.uniforms(400, (32, 3)) * 2.2 - 1.1)
    arrs["pe_x"] = x
    arrs["pe_10"] = ref_model.positional_encoding(x, dim=10)
    arrs["pe_4"] = ref_model.positional_encoding(x, dim=4)
    arrs["pe_kat"] = ref_model.positional_encoding(torch.Tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]]), dim=1)
    for kind, seed in (("init", 1), ("dense", 2)):
        sd = synthetic.make_state_dict(seed, kind)
        net = load_model(sd)
        pts = T(synthetic.uniforms(410 + seed, (8, 16, 3)) * 6.0 - 3.0)
        dr = T(synthetic.uniforms(420 + seed, (8, 3)) * 2.0 - 1.0)
        with torch.no_grad():
            sg, rgb = net.fine_network(pts, dr)
        arrs.update({f"pts_{kind}": pts, f"dir_{kind}": dr, f"sigma_{kind}": sg, f"rgb_{kind}": rgb})
    save("mlp", **arrs)


def gen_network():
    arrs = {}
    N = 64
    o, d, xs, ys, c2w, focal = orbit_rays(45.0, N, 51)
    target = T(synthetic.uniforms(53, (N, 3)))
    arrs.update(o=o, d=d, xs=xs, ys=ys, c2w=c2w, focal=np.float64(focal), target=target)
    small = ("bias", "density_fn.0.weight", "rgb_fn.2.weight")
    for kind, seed in (("init", 3), ("dense", 4)):
        sd = synthetic.make_state_dict(seed, kind)
        net = load_model(sd)
        cap = {}
        hooks = [net.coarse_network.register_forward_hook(
                    lambda m, i, out: cap.update(c_pts=i[0], c_sigma=out[0], c_rgb=out[1])),
                 net.fine_network.register_forward_hook(
                    lambda m, i, out: cap.update(f_in_pts=i[0], f_sigma=out[0], f_rgb=out[1]))]
        real_sort, real_ss = torch.sort, torch.searchsorted
        torch.sort = lambda *a, **k: cap.setdefault("sorted", real_sort(*a, **k))
        torch.searchsorted = lambda *a, **k: cap.setdefault("idx", real_ss(*a, **k))
        try:
            with RandStream(500 + seed * 10):
                batch = {"origin": o[None].clone(), "direc": d[None].clone(), "rgb": target[None].clone()}
                loss = net.training_step(batch, 0)
        finally:
            torch.sort, torch.searchsorted = real_sort, real_ss
            for h in hooks:
                h.remove()
        loss.backward()
        with RandStream(500 + seed * 10), torch.no_grad():
            pred = net.forward(o, d)
        arrs.update({f"loss_{kind}": loss.detach(), f"fine_rgb_rays_{kind}": pred["fine_rgb_rays"],
                     f"coarse_rgb_rays_{kind}": pred["coarse_rgb_rays"],
                     f"sorted_ts_{kind}": cap["sorted"][0].detach(), f"idx_{kind}": cap["idx"],
                     f"c_sigma_{kind}": cap["c_sigma"].detach(), f"c_rgb_{kind}": cap["c_rgb"].detach(),
                     f"f_in_pts_{kind}": cap["f_in_pts"].detach(),
                     f"f_sigma_{kind}": cap["f_sigma"].detach(), f"f_rgb_{kind}": cap["f_rgb"].detach(),
                     f"stat_{kind}": np.array([float(net.logged[k]) for k in (
                         "coarse_density_norms", "coarse_density_non_zeros",
                         "fine_density_norms", "fine_density_non_zeros")])})
        names, norms = [], []
        for k, p in net.named_parameters():
            names.append(k)
            norms.append(float(p.grad.norm()))
            if any(k.endswith(s) for s in small):
                arrs[f"grad_{kind}__{k}"] = p.grad.clone()
            else:
                arrs[f"gradhead_{kind}__{k}"] = p.grad[:4, :8].clone()
        arrs[f"grad_names_{kind}"] = np.array(names)
        arrs[f"grad_norms_{kind}"] = np.array(norms)
    save("network", **arrs)


def gen_render():
    sd = synthetic.make_state_dict(5, "dense")
    net = load_model(sd)
    H = W = 100
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    c2w = ref_helpers.pose_spherical(30.0, -30.0, 4.0)
    o, d = ref_data.get_rays(H, W, focal, c2w)
    with RandStream(600):
        im = ref_helpers.view_reconstruction(net, o, d, N=4096)
    save("render100", image=im, c2w=c2w, focal=np.float64(focal))


TRAJ_STEPS, TRAJ_RAYS, TRAJ_VIEWS = 200, 512, 8


def trajectory_batch(step, images=None):
    """Batch `step` of the training-trajectory fixture, a pure function of the step (tests/test_gpu_trajectory.py rebuilds it on
    the device): TRAJ_RAYS centre-cropped pixels (dataloader.py:13-34 with cropping) of orbit view step % TRAJ_VIEWS of the analytic
    scene, colours from that view's image."""
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    v = step % TRAJ_VIEWS
    c2w = synthetic.orbit_pose(-180.0 + 45.0 * v, -30.0, 4.0)
    xs = (synthetic.uniforms(9000 + 2 * step, (TRAJ_RAYS,)) * (W // 2)).astype(np.int64) + W // 4
    ys = (synthetic.uniforms(9001 + 2 * step, (TRAJ_RAYS,)) * (H // 2)).astype(np.int64) + H // 4
    return v, c2w, focal, xs, ys


def gen_trajectory():
    """~200 optimiser steps of the UNMODIFIED reference (NeRFNetwork.training_step + the Adam of configure_optimizers,
    nerf_model.py:134-169) from the synthetic random-init weights on recorded batches / uniforms: the loss trajectory, the
    "briefly-trained" weight set it ends on, a 1024-ray step on that set with all 40 gradients (8c(vi)) and
    a 100x100 frame rendered with it next to the analytic ground truth (delta-PSNR fixture)."""
    import time
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    images = []
    for v in range(TRAJ_VIEWS):
        c2w = synthetic.orbit_pose(-180.0 + 45.0 * v, -30.0, 4.0)
        images.append(synthetic.analytic_scene_rgba(c2w.numpy(), H, W, focal)[..., :3].astype(np.float32) / 255.0)
    net = load_model(synthetic.make_state_dict(0, "init"))
    opt = net.configure_optimizers()["optimizer"]
    assert type(opt).__name__ == "Adam" and opt.param_groups[0]["lr"] == 5e-4
    losses, stats = [], []
    t0 = time.time()
    for step in range(TRAJ_STEPS):
        v, c2w, focal, xs, ys = trajectory_batch(step)
        o, d = ref_data.get_rays(H, W, focal, c2w)
        o, d = o[ys, xs, :].contiguous(), d[ys, xs, :].contiguous()
        rgb = T(images[v][ys, xs, :])
        with RandStream(20000 + 3 * step):
            loss = net.training_step({"origin": o[None].clone(), "direc": d[None].clone(), "rgb": rgb[None].clone()}, step)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append([float(net.logged["train_loss"]), float(net.logged["train_coarse_loss"]), float(net.logged["train_fine_loss"])])
        stats.append([float(net.logged[k]) for k in ("coarse_density_norms", "coarse_density_non_zeros", "fine_density_norms", "fine_density_non_zeros")])
        if step % 10 == 0:
            print(f"step {step}: loss {losses[-1][0]:.5f} fine nz {stats[-1][3]:.0f} ({time.time() - t0:.0f} s)", flush=True)
    arrs = {"losses": np.array(losses), "stats": np.array(stats), "steps": np.int64(TRAJ_STEPS), "rays": np.int64(TRAJ_RAYS)}
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        arrs[f"sd__{k}"] = v
    # ---- a 1024-ray training step on the trained set: loss, forward, every gradient (full tensors for the small ones,
    # 16 x 32 corner blocks + norms for the 256-wide ones)
    N = 1024
    o, d, xs, ys, c2w, focal = orbit_rays(70.0, N, 71)
    xs, ys = xs // 2 + 200, ys // 2 + 200
    o_all, d_all = ref_data.get_rays(H, W, focal, c2w)
    o, d = o_all[ys, xs, :].contiguous(), d_all[ys, xs, :].contiguous()
    img = synthetic.analytic_scene_rgba(c2w.numpy(), H, W, focal)[..., :3].astype(np.float32) / 255.0
    target = T(img[ys.numpy(), xs.numpy(), :])
    net.zero_grad()
    with RandStream(30000):
        loss = net.training_step({"origin": o[None].clone(), "direc": d[None].clone(), "rgb": target[None].clone()}, 0)
    loss.backward()
    with RandStream(30000), torch.no_grad():
        pred = net.forward(o, d)
    arrs.update(g_o=o, g_d=d, g_target=target, g_loss=loss.detach(), g_fine=pred["fine_rgb_rays"], g_coarse=pred["coarse_rgb_rays"])
    names, norms = [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(float(p.grad.norm()))
        arrs[f"grad__{k}"] = p.grad.clone() if p.grad.numel() <= 1024 else p.grad[:16, :32].clone()
    arrs["grad_names"] = np.array(names)
    arrs["grad_norms"] = np.array(norms)
    # ---- 100 x 100 frame with the trained set + the analytic ground truth of the same camera
    Hs = Ws = 100
    focal_s = 0.5 * Ws / np.tan(0.5 * 0.6911112070083618)
    c2w = ref_helpers.pose_spherical(30.0, -30.0, 4.0)
    o, d = ref_data.get_rays(Hs, Ws, focal_s, c2w)
    with RandStream(31000):
        im = ref_helpers.view_reconstruction(net, o, d, N=4096)
    arrs.update(frame=im, frame_gt=synthetic.analytic_scene_rgba(c2w.numpy(), Hs, Ws, focal_s)[..., :3], frame_c2w=c2w,
                frame_focal=np.float64(focal_s))
    save("trajectory", **arrs)


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:] or ["rays", "coarse", "composite", "fine", "pe_mlp", "network", "render"]
    for w in which:
        globals()[f"gen_{w}"]()
