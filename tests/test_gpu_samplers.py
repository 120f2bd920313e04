"""GPU parity: the HBM-bound kernels (through the C-ABI via the host modules) against the CPU oracle and the
golden vectors.  Depths, points and bin indices must be bit-exact; exp-based quantities within 1e-6."""
import hashlib

import numpy as np
import pytest
import torch

import synthetic
from oracle import nerf_oracle as O
from util import T, bits_equal, rand_triple

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def mods():
    import dataloader
    import nerf_helpers
    import nerf_model
    return dataloader, nerf_helpers, nerf_model


def test_raygen(golden, mods):
    dl, _, _ = mods
    g = golden["rays"]
    focal800 = float(g["focal800"])
    for k in range(4):
        pose = T(g["poses"][k])
        o, d = dl.get_rays(48, 64, 77.25, pose)
        assert o.is_cuda and d.shape == (48, 64, 3)
        assert bits_equal(d, g[f"d_small_{k}"]) and bits_equal(o, g[f"o_small_{k}"])
        o, d = dl.get_rays(800, 800, focal800, pose)
        assert hashlib.sha256((d.cpu().numpy() + 0.0).tobytes()).hexdigest() == str(g["d800_sha256"][k])
    n = golden["network"]
    o, d = dl.get_rays_at(800, 800, float(n["focal"]), T(n["c2w"]), T(n["xs"], DEV), T(n["ys"], DEV))
    assert bits_equal(o, n["o"]) and bits_equal(d, n["d"])
    # ragged sizes: the kernel takes four rays per thread with 16-byte stores and finishes the remainder one ray per thread
    pose = T(g["poses"][1])
    for (Hh, Ww) in ((7, 9), (1, 1), (3, 1), (5, 5)):
        o, d = dl.get_rays(Hh, Ww, 31.5, pose)
        ro, rd = O.get_rays(Hh, Ww, 31.5, pose)
        assert bits_equal(d, rd) and bits_equal(o, ro.expand(Hh, Ww, 3).contiguous()), (Hh, Ww)
    xs, ys = T(n["xs"], DEV)[:1023], T(n["ys"], DEV)[:1023]
    o, d = dl.get_rays_at(800, 800, float(n["focal"]), T(n["c2w"]), xs, ys)
    assert bits_equal(o, n["o"][:1023]) and bits_equal(d, n["d"][:1023])


def test_coarse_samples(golden, mods):
    _, h, _ = mods
    g = golden["coarse"]
    o, d = T(g["o"], DEV), T(g["d"], DEV)
    pts, ts = h.generate_coarse_samples(o, d, 64, 2.0, 6.0, rand=T(synthetic.uniforms(100, (256, 64)), DEV))
    assert pts.shape == (256, 64, 3) and ts.shape == (256, 64, 1)
    assert bits_equal(pts, g["pts"]) and bits_equal(ts, g["ts"])
    pts, ts = h.generate_coarse_samples(o[:5], d[:5], 7, 0.3, 5.1, rand=T(synthetic.uniforms(110, (5, 7)), DEV))
    assert bits_equal(pts, g["pts_odd"]) and bits_equal(ts, g["ts_odd"])
    # reference test (tests/nerf_helpers_test.py:49-63), inputs moved to the device, internal torch.rand draw
    pts, ts = h.generate_coarse_samples(torch.zeros(1, 3, device=DEV), torch.ones(1, 3, device=DEV), 2)
    lo, hi = torch.tensor([2.0, 4.0], device=DEV), torch.tensor([4.0, 6.0], device=DEV)
    assert ((lo <= ts[0, :, 0]) & (ts[0, :, 0] < hi)).all()
    # empty batch
    pts, ts = h.generate_coarse_samples(torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), 64)
    assert pts.shape == (0, 64, 3)


def test_compositing(golden, mods):
    _, h, _ = mods
    g = golden["composite"]
    for S in (64, 192):
        ts, sg, rgb = T(g[f"ts_{S}"], DEV), T(g[f"sigma_{S}"], DEV), T(g[f"rgb_{S}"], DEV)
        dl = h.generate_deltas(ts)
        assert bits_equal(dl, g[f"deltas_{S}"])
        w = h.calculate_unnormalized_weights(sg, dl)
        torch.testing.assert_close(w.cpu(), T(g[f"weights_{S}"]), atol=1e-6, rtol=0)
        col = h.estimate_ray_color(w, rgb)
        torch.testing.assert_close(col.cpu(), T(g[f"color_{S}"]), atol=2e-6, rtol=1e-6)
        c = h.composite(sg, rgb, ts)
        torch.testing.assert_close(c["weights"].cpu(), T(g[f"weights_{S}"]), atol=1e-6, rtol=0)
        torch.testing.assert_close(c["rgb"].cpu(), T(g[f"color_{S}"]), atol=2e-6, rtol=1e-6)
        dep, acc = O.depth_and_acc(T(g[f"weights_{S}"]), T(g[f"ts_{S}"]))
        torch.testing.assert_close(c["depth"].cpu(), dep[:, 0], atol=1e-5, rtol=1e-6)
        torch.testing.assert_close(c["acc"].cpu(), acc[:, 0], atol=2e-6, rtol=1e-6)
        sgc = T(g[f"sigma_{S}"])
        torch.testing.assert_close(c["stats"].cpu(), torch.stack([(sgc ** 2).sum(), (sgc != 0).sum().float()]), rtol=1e-5, atol=0)
        # all-zero density rays render black with zero opacity
        empty = (sg == 0).all(dim=1)[:, 0]
        assert empty.sum() >= 8 and (c["rgb"][empty] == 0).all() and (c["acc"][empty] == 0).all()


def test_compositing_reference_kats(mods):
    _, h, _ = mods
    w = h.calculate_unnormalized_weights(torch.tensor([0, 50, 1, 0.3, 1.0], device=DEV).view(1, 5, 1), torch.full((1, 5, 1), 0.2, device=DEV))
    torch.testing.assert_close(w.cpu(), torch.tensor([0, 0.9999546001, 8.229611e-6, 2.1646e-6, 6.34545e-6]).view(1, 5, 1))
    dl = h.generate_deltas(torch.arange(2, 6, 1, device=DEV).view(1, -1, 1))          # int64 input, as upstream's test
    torch.testing.assert_close(dl.cpu(), torch.tensor([1, 1, 1, 1e10]).view(1, 4, 1))
    col = h.estimate_ray_color(torch.full((1, 256, 1), 1 / 256, device=DEV), torch.ones(1, 256, 3, device=DEV))
    torch.testing.assert_close(col.cpu(), torch.ones(1, 3))
    w = torch.zeros(1, 256, 1, device=DEV); w[:, 200] = 1.0
    torch.testing.assert_close(h.estimate_ray_color(w, torch.ones(1, 256, 3, device=DEV)).cpu(), torch.ones(1, 3))


def test_fine_samples(golden, mods):
    _, h, _ = mods
    g = golden["fine"]
    o, d = T(g["o"], DEV), T(g["d"], DEV)
    rand = (T(synthetic.uniforms(320, (256, 1)), DEV), T(synthetic.uniforms(321, (256, 128, 1)), DEV))
    pts, fts, idx = h.inverse_transform_sampling(o, d, T(g["w"], DEV), T(g["c_ts"], DEV), 128, rand=rand, return_idx=True)
    assert pts.shape == (256, 128, 3) and fts.shape == (256, 128, 1)
    assert bits_equal(idx, g["idx"]) and bits_equal(fts, g["f_ts"]) and bits_equal(pts, g["f_pts"])
    rand = (T(synthetic.uniforms(332, (9, 1)), DEV), T(synthetic.uniforms(333, (9, 5, 1)), DEV))
    pts, fts = h.inverse_transform_sampling(o[:9], d[:9], T(g["w_odd"], DEV), T(g["c_ts_odd"], DEV), 5, rand=rand)
    assert bits_equal(fts, g["f_ts_odd"]) and bits_equal(pts, g["f_pts_odd"])


def test_merge_sort(golden, mods):
    _, h, _ = mods
    g, n = golden["fine"], golden["network"]
    o, d = T(g["o"], DEV), T(g["d"], DEV)
    pts, ts = h.merge_samples(o, d, T(g["f_ts"], DEV), T(g["c_ts"], DEV))
    ref_pts, ref_ts = O.merge_sorted(T(g["f_pts"]), T(g["f_ts"]), T(g["c_pts"]), T(g["c_ts"]))
    assert bits_equal(ts, ref_ts) and bits_equal(pts, ref_pts)
    assert (ts[:, 1:, 0] >= ts[:, :-1, 0]).all()
    # ties and reversed input
    a = torch.tensor([[5.0, 3.0, 3.0, 1.0]], device=DEV).view(1, 4, 1)
    b = torch.tensor([[3.0, 0.5]], device=DEV).view(1, 2, 1)
    _, ts = h.merge_samples(torch.zeros(1, 3, device=DEV), torch.ones(1, 3, device=DEV), a, b)
    assert ts.flatten().tolist() == [0.5, 1.0, 3.0, 3.0, 3.0, 5.0]
    # general path (A + B > 256, enumeration sort in shared memory) and odd sizes of the register bitonic path
    for A, B in ((300, 77), (5, 2), (255, 1), (1, 31)):
        ta = torch.rand(33, A, 1, device=DEV) * 4 + 2
        tb = torch.sort(torch.rand(33, B, 1, device=DEV) * 4 + 2, dim=1).values
        ta[:, :: 3] = tb[:, :1]                                           # plenty of exact ties
        _, ts = h.merge_samples(torch.zeros(33, 3, device=DEV), torch.ones(33, 3, device=DEV), ta, tb, want_points=False)
        assert torch.equal(ts, torch.sort(torch.cat([ta, tb], 1), dim=1).values)


def test_sampler_properties_full_size(mods):
    """BASELINE config sizes (4096 rays, 64 + 128): size-independent properties instead of an oracle run."""
    _, h, _ = mods
    N = 4096
    o = torch.zeros(N, 3, device=DEV)
    d = torch.nn.functional.normalize(torch.randn(N, 3, device=DEV), dim=1) * 1.05
    pts, ts = h.generate_coarse_samples(o, d, 64)
    t = ts[..., 0]
    edges = torch.arange(2.0, 6.0, 4.0 / 64, device=DEV)
    assert ((t >= edges[None]) & (t < edges[None] + 4.0 / 64 + 1e-6)).all()
    sigma = torch.relu(torch.randn(N, 64, 1, device=DEV) * 3 - 2)
    sigma[::3] = 0
    c = h.composite(sigma, torch.rand(N, 64, 3, device=DEV), ts)
    w = c["weights"]
    assert (w >= 0).all() and (c["acc"] <= 1 + 1e-5).all()
    fp, ft, idx = h.inverse_transform_sampling(o, d, w, ts, 128, return_idx=True)
    assert (idx >= 0).all() and (idx <= 64).all() and (idx[:, 1:] >= idx[:, :-1]).all()
    assert (idx[::3] == 64).all()                                       # NaN cdf rows: last bin
    assert (ft >= 2.0).all() and (ft <= 6.0).all()
    _, ts_all = h.merge_samples(o, d, ft, ts, want_points=False)
    assert (ts_all[:, 1:] >= ts_all[:, :-1]).all()
    both, _ = torch.sort(torch.cat([ft, ts], 1), 1)
    assert torch.equal(ts_all, both)                                    # a permutation of the inputs: sortedness + multiset


def test_positional_encoding(golden, mods):
    _, _, m = mods
    g = golden["mlp"]
    x = T(g["pe_x"], DEV)
    torch.testing.assert_close(m.positional_encoding(x, 10).cpu(), T(g["pe_10"]), atol=2e-6, rtol=0)
    torch.testing.assert_close(m.positional_encoding(x, 4).cpu(), T(g["pe_4"]), atol=2e-6, rtol=0)
    kat = m.positional_encoding(torch.tensor([[0.0, 0, 0], [1.0, 1, 1]], device=DEV), 1).cpu()     # nerf_model_test.py:41-58
    torch.testing.assert_close(kat, torch.tensor([[1.0, 1, 1, 0, 0, 0], [-1.0, -1, -1, 0, 0, 0]]), atol=1e-6, rtol=0)
    assert m.positional_encoding(torch.rand(4096, 64, 3, device=DEV), 10).shape == (4096, 64, 60)   # nerf_model_test.py:60-63


def test_mlp_fp32(golden, mods):
    _, _, m = mods
    g = golden["mlp"]
    for kind, seed in (("init", 1), ("dense", 2)):
        net = m.NeRFNetwork(precision="fp32")
        net.load_state_dict(synthetic.make_state_dict(seed, kind))
        net = net.to(DEV)
        sg, rgb = net.fine_network(T(g[f"pts_{kind}"], DEV), T(g[f"dir_{kind}"], DEV))
        assert sg.shape == (8, 16, 1) and rgb.shape == (8, 16, 3)
        torch.testing.assert_close(sg.cpu(), T(g[f"sigma_{kind}"]), atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(rgb.cpu(), T(g[f"rgb_{kind}"]), atol=2e-5, rtol=1e-4)
    # the reference's shape test (nerf_model_test.py:69-72) with non-default encoding sizes
    small = m.NeRFModel(position_dim=6, direction_dim=2, precision="fp32").to(DEV)
    sg, rgb = small(torch.rand(4, 4, 3, device=DEV), torch.rand(4, 3, device=DEV))
    assert sg.shape == (4, 4, 1) and rgb.shape == (4, 4, 3) and (sg >= 0).all() and ((rgb > 0) & (rgb < 1)).all()


def test_compositing_large_batch_kernel(mods):
    """nerf_composite switches to the thread-per-ray kernel (tiles staged through shared memory) from 32768 rays on: the
    running sum is taken in sample order there and by a 5-step shuffle scan in the warp-per-ray kernel, the per-ray sums in
    sample order vs lane partials + tree: a 40000-ray batch must match its two halves to a few ulp (2e-6) in weights,
    colour, depth and opacity, and the density statistics to rounding."""
    _, h, _ = mods
    N = 40000
    for S in (64, 192, 37):
        gen = torch.Generator(device=DEV).manual_seed(100 + S)
        ts = torch.sort(2 + 4 * torch.rand(N, S, 1, device=DEV, generator=gen), dim=1).values.contiguous()
        sg = torch.relu(torch.randn(N, S, 1, device=DEV, generator=gen) * 3 - 2)
        sg[::7] = 0
        rgb = torch.rand(N, S, 3, device=DEV, generator=gen)
        full = h.composite(sg, rgb, ts)
        halves = [h.composite(sg[a:b].contiguous(), rgb[a:b].contiguous(), ts[a:b].contiguous()) for a, b in ((0, N // 2), (N // 2, N))]
        for key in ("weights", "rgb", "depth", "acc"):
            torch.testing.assert_close(full[key], torch.cat([x[key] for x in halves]), atol=2e-6, rtol=2e-6)
        torch.testing.assert_close(full["stats"], halves[0]["stats"] + halves[1]["stats"], rtol=1e-5, atol=0)
        torch.testing.assert_close(full["norm"], torch.sqrt(full["stats"][0]), rtol=1e-6, atol=0)
