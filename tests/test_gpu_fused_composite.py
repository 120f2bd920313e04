"""GPU parity of the compositing fused into the tensor-core MLP kernel (nerf_mlp_composite_tc, mlp_tc3.cu COMP form)
against the two-launch path it replaces (nerf_mlp_forward_tc[_train] + nerf_composite), which is itself checked against
the oracle in test_gpu_mlp_tc.py / test_gpu_samplers.py.  Both spell the same arithmetic in the same order, so every
output is compared BIT FOR BIT; only the two atomically accumulated density statistics get a relative tolerance."""
import pytest
import torch

import synthetic
from util import T, bits_equal, rand_triple

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_net(seed=4, kind="dense"):
    import nerf_model
    net = nerf_model.NeRFNetwork(precision="bf16")
    net.load_state_dict(synthetic.make_state_dict(seed, kind))
    return net.to(DEV)


def rays(N, S, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    o = torch.randn(N, 3, device=DEV, generator=g) * 0.3
    d = torch.nn.functional.normalize(torch.randn(N, 3, device=DEV, generator=g), dim=1) * 1.07
    ts = torch.sort(2.0 + 4.0 * torch.rand(N, S, 1, device=DEV, generator=g), dim=1).values.contiguous()
    return o, d, ts


# ray counts: many pairs per CTA / fewer groups than SMs / odd (partial last group) / single ray;
# sample counts: the two the network uses (64, 192) + every other group shape (1, 2 and 4 rays per group; 1..5 tiles)
@pytest.mark.parametrize("N,S", [(4096, 64), (4096, 192), (1024, 192), (301, 192), (301, 64), (1, 192), (1, 64), (2, 64),
                                 (515, 128), (130, 256), (257, 32), (203, 96), (99, 160), (51, 320)])
def test_fused_matches_two_launch_path(N, S):
    import nerf_helpers as h
    net = make_net().fine_network
    assert net.can_composite(S)
    o, d, ts = rays(N, S, 10 * S + N)
    sg, rgb = net.forward_rays(o, d, ts)
    ref = h.composite(sg, rgb, ts)
    got = net.render_rays(o, d, ts, want_weights=True, keep_samples=True)
    torch.cuda.synchronize()
    assert bits_equal(got["sigma"], sg) and bits_equal(got["rgb_samples"], rgb)
    for key in ("weights", "rgb", "depth", "acc"):
        assert bits_equal(got[key], ref[key]), key
    torch.testing.assert_close(got["stats"], ref["stats"], rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(got["norm"], ref["norm"], rtol=1e-5, atol=1e-3)
    # render form: no per-sample outputs, no weights
    lean = net.render_rays(o, d, ts, want_weights=False)
    torch.cuda.synchronize()
    assert lean["sigma"] is None and lean["weights"] is None
    for key in ("rgb", "depth", "acc"):
        assert bits_equal(lean[key], ref[key]), key


# (2, 64), (1, 192): a single ray group runs as one plain CTA (no 2-CTA cluster / multicast) - the other launch form of the kernel
@pytest.mark.parametrize("N,S", [(1024, 192), (301, 64), (2, 64), (1, 192)])
def test_fused_training_form_saves_the_same_tensors(N, S):
    import nerf_helpers as h
    import training
    net = make_net().coarse_network
    o, d, ts = rays(N, S, 77 + N)
    sg, rgb, (acts, masks) = training.mlp_forward_train(net, o, d, ts)
    ref = h.composite(sg, rgb, ts)
    got = net.render_rays(o, d, ts, want_weights=True, save=True)
    torch.cuda.synchronize()
    assert bits_equal(got["sigma"], sg) and bits_equal(got["rgb_samples"], rgb)
    M = N * S
    a0, a1 = training.untile(acts, M, training.ACT), training.untile(got["saved"][0], M, training.ACT)
    assert torch.equal(a0.view(torch.int16), a1.view(torch.int16))
    words = training.ACT // 64
    tiles = training.padded_rows(M) // 128
    m0 = masks.view(tiles, words, 128).permute(0, 2, 1).reshape(-1, words)[:M]
    m1 = got["saved"][1].view(tiles, words, 128).permute(0, 2, 1).reshape(-1, words)[:M]
    assert torch.equal(m0, m1)
    for key in ("weights", "rgb", "depth", "acc"):
        assert bits_equal(got[key], ref[key]), key


def test_unsupported_sample_counts_are_rejected():
    import _native as nat
    lib = nat.lib()
    assert [S for S in (32, 64, 96, 128, 160, 192, 256, 320) if not lib.nerf_mlp_composite_tc_supported(S)] == []
    assert [S for S in (0, 5, 40, 77, 224, 448) if lib.nerf_mlp_composite_tc_supported(S)] == []
    net = make_net().fine_network
    o, d, ts = rays(8, 40, 1)
    with pytest.raises(RuntimeError, match="not supported"):
        net.render_rays(o, d, ts)


def test_network_forward_identical_with_and_without_fusion(golden):
    """NeRFNetwork.forward end to end (coarse -> fine sampling -> fine): fused and two-launch paths agree bit for bit,
    so every oracle tolerance established for one holds for the other."""
    import training
    g = golden["network"]
    net = make_net(4, "dense")
    o, d = T(g["o"], DEV), T(g["d"], DEV)
    rand = rand_triple(540, 64, device=DEV)
    net.keep_samples = True
    try:
        training.FUSE_COMPOSITE = False
        a = net.forward(o, d, rand=rand)
        last_a = dict(net.last)
        training.FUSE_COMPOSITE = True
        b = net.forward(o, d, rand=rand)
        last_b = dict(net.last)
    finally:
        training.FUSE_COMPOSITE = True
    torch.cuda.synchronize()
    for key in ("coarse_rgb_rays", "fine_rgb_rays"):
        assert bits_equal(a[key], b[key]), key
    for key in ("depth", "acc", "ts", "coarse_weights", "coarse_sigma", "fine_sigma", "fine_rgb"):
        assert bits_equal(last_a[key], last_b[key]), key


def test_strata_inside_the_kernel_are_bit_identical(golden):
    """nerf_mlp_composite_tc_strata forms the stratified depths in the coarse kernel's producer warps: depths, weights and every
    output of NeRFNetwork.forward are bit for bit those of nerf_coarse_sample + nerf_mlp_composite_tc."""
    import nerf_helpers as h
    import training
    net = make_net(4, "dense")
    for N in (4096, 301, 2):
        o, d, _ = rays(N, 64, 40 + N)
        u = torch.rand(N, 64, device=DEV)
        t_base, step = h._strata(2.0, 6.0, 64, o.device)
        _, ts = h.generate_coarse_samples(o, d, 64, rand=u)
        ref = net.coarse_network.render_rays(o, d, ts, keep_samples=True)
        got = net.coarse_network.render_rays(o, d, None, keep_samples=True, strata=(u, t_base, step))
        torch.cuda.synchronize()
        assert bits_equal(got["ts"], ts)
        for key in ("weights", "rgb", "depth", "acc", "sigma", "rgb_samples"):
            assert bits_equal(got[key], ref[key]), (N, key)
    g = golden["network"]
    o, d = T(g["o"], DEV), T(g["d"], DEV)
    rand = rand_triple(540, 64, device=DEV)
    outs = []
    try:
        for flag in (False, True):
            training.FUSE_STRATA = flag
            out = net.forward(o, d, rand=rand)
            outs.append((out["coarse_rgb_rays"].clone(), out["fine_rgb_rays"].clone(), net.last["ts"].clone(), net.last["coarse_ts"].clone()))
    finally:
        training.FUSE_STRATA = True
    torch.cuda.synchronize()
    assert all(bits_equal(a, b) for a, b in zip(*outs))


def test_training_gradients_identical_with_and_without_fusion():
    import training
    net = make_net(3, "init")
    o, d, _ = rays(512, 64, 5)
    target = torch.rand(512, 3, device=DEV)
    rand = rand_triple(900, 512, device=DEV)
    grads = []
    try:
        for fuse in (False, True):
            training.FUSE_COMPOSITE = fuse
            net.zero_grad(set_to_none=True)
            out = net.forward(o, d, rand=rand)
            loss = ((out["coarse_rgb_rays"] - target) ** 2).mean() + ((out["fine_rgb_rays"] - target) ** 2).mean()
            loss.backward()
            grads.append([p.grad.clone() for p in net.parameters()])
    finally:
        training.FUSE_COMPOSITE = True
    torch.cuda.synchronize()
    for ga, gb in zip(*grads):
        # wgrad accumulates with atomics: equal up to summation order
        torch.testing.assert_close(ga, gb, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("N,C,F", [(4096, 64, 128), (301, 64, 128), (77, 5, 7), (33, 100, 156), (9, 2, 3), (1, 64, 128)])
def test_fine_depths_sorted_matches_sampler_plus_merge(N, C, F):
    """nerf_fine_sample_merge (K3 + K4 in one launch) against nerf_fine_sample + nerf_merge_sort: bit-identical sorted
    depths, including rays whose weights are all zero (NaN cdf, nerf_helpers.py:138)."""
    import nerf_helpers as h
    g = torch.Generator(device=DEV).manual_seed(N + C)
    o, d, _ = rays(N, 32, 3)
    _, c_ts = h.generate_coarse_samples(o, d, C, rand=torch.rand(N, C, device=DEV, generator=g))
    w = torch.rand(N, C, 1, device=DEV, generator=g) ** 4
    w[::7] = 0.0                                                  # dead rays
    if N > 3:
        w[3, : C // 2] = 0.0                                      # flat cdf prefix (ties in the search)
    rand = (torch.rand(N, 1, device=DEV, generator=g), torch.rand(N, F, 1, device=DEV, generator=g))
    rand[1][::5, ::3] = 0.0                                       # jitter 0: the fine depth EQUALS a coarse depth (ties in the rank merge)
    _, f_ts = h.inverse_transform_sampling(o, d, w, c_ts, F, rand=rand)
    _, ref = h.merge_samples(o, d, f_ts, c_ts, want_points=False)
    got = h.fine_depths_sorted(w, c_ts, F, rand=rand)
    torch.cuda.synchronize()
    assert got.shape == (N, C + F, 1) and bits_equal(got, ref)
    assert bool((got[:, 1:] >= got[:, :-1]).all())
    # a caller's own, UNSORTED coarse depths (every second ray reversed): those rays take the general 256-wide network
    if C > 1:
        c_mix = c_ts.clone()
        c_mix[1::2] = c_mix[1::2].flip(1)
        _, f_ts = h.inverse_transform_sampling(o, d, w, c_mix, F, rand=rand)
        _, ref = h.merge_samples(o, d, f_ts, c_mix, want_points=False)
        got = h.fine_depths_sorted(w, c_mix, F, rand=rand)
        torch.cuda.synchronize()
        assert bits_equal(got, ref)


@pytest.mark.parametrize("N", [1, 2, 301, 4096, 20000])
def test_render_forward_one_call_matches_three_calls(N):
    """nerf_render_forward (coarse kernel, sampler, fine kernel queued by ONE call of the C ABI, intermediates in a workspace)
    against the same three launches issued by the Python host: every output bit-identical, statistics included."""
    import training
    net = make_net(5, "dense")
    o, d, _ = rays(N, 32, 9 + N)
    rand = rand_triple(700 + N, N, device=DEV)
    outs = []
    try:
        for flag in (True, False):
            training.RENDER_ONE_CALL = flag
            with torch.no_grad():
                out = net.forward(o, d, rand=rand)
            torch.cuda.synchronize()
            outs.append([out["fine_rgb_rays"].clone(), out["coarse_rgb_rays"].clone(), net.last["ts"].clone(), net.last["coarse_ts"].clone(),
                         net.last["coarse_weights"].clone(), net.last["depth"].clone(), net.last["acc"].clone(),
                         torch.stack([net.logged[k].clone() for k in ("coarse_density_non_zeros", "fine_density_non_zeros")])])
    finally:
        training.RENDER_ONE_CALL = True
    for a, b in zip(outs[0][:-1], outs[1][:-1]):
        assert bits_equal(a, b)
    torch.testing.assert_close(outs[0][-1], outs[1][-1])


def test_full_frame_is_independent_of_chunking_and_sharding():
    """BASELINE configs[1] size (800 x 800 = 640 000 rays, 64 + 128 samples): with the uniforms fixed per ray, the image must not
    depend on how the rays are cut into chunks (4096 as the reference, an odd 4095, 12 345, all 640 000 at once) or into per-GPU slabs - every ray is
    composited by exactly one CTA from its own samples, whatever tile range that CTA was given.  Also: opacity <= 1, expected
    depth inside [near, far] where the ray hit anything, sorted merged depths."""
    import dataloader
    import multi_gpu
    import nerf_helpers as h
    net = make_net(5, "dense")
    H = W = 800
    n = H * W
    focal = 0.5 * W / 0.36                                    # ~ the lego camera
    o, d = dataloader.get_rays(H, W, focal, h.pose_spherical(35.0, -30.0, 4.0), device=DEV)
    o, d = o.reshape(n, 3), d.reshape(n, 3)
    g = torch.Generator(device=DEV).manual_seed(11)
    u_c = torch.rand(n, 64, device=DEV, generator=g)
    eps = torch.rand(n, 1, device=DEV, generator=g)
    u_f = torch.rand(n, 128, 1, device=DEV, generator=g)

    def render(chunk, lo=0, hi=n):
        out = torch.empty((hi - lo, 3), device=DEV)
        depth = torch.empty((hi - lo,), device=DEV)
        acc = torch.empty((hi - lo,), device=DEV)
        with torch.no_grad():
            for i in range(lo, hi, chunk):
                j = min(i + chunk, hi)
                out[i - lo:j - lo] = net.forward(o[i:j], d[i:j], rand=(u_c[i:j], eps[i:j], u_f[i:j]))["fine_rgb_rays"]
                depth[i - lo:j - lo], acc[i - lo:j - lo] = net.last["depth"], net.last["acc"]
                if i == lo:
                    ts = net.last["ts"]
                    assert bool((ts[:, 1:] >= ts[:, :-1]).all())
        return out, depth, acc

    ref, depth, acc = render(4096)
    assert torch.isfinite(ref).all() and float(acc.max()) <= 1.0 + 1e-5 and float(acc.min()) >= 0.0
    hit = acc > 0.5
    assert bool(hit.any())
    mean_depth = depth[hit] / acc[hit]
    assert float(mean_depth.min()) >= 2.0 - 1e-3 and float(mean_depth.max()) <= 6.0 + 1e-3
    for chunk in (4095, 12345, n):                           # n: the whole frame through one set of launches
        got, _, _ = render(chunk)
        assert bits_equal(got, ref), chunk
    for ws in (3, 8):                                        # per-GPU slabs of multi_gpu.ray_slab, each rendered on its own
        for rank in (0, ws - 1):
            lo, hi = multi_gpu.ray_slab(n, rank, ws)
            got, _, _ = render(4096, lo, hi)
            assert bits_equal(got, ref[lo:hi]), (ws, rank)
    torch.cuda.synchronize()


def test_grouped_launches_render_the_same_scene():
    """render_rays_chunked groups the reference's N-ray chunks into launches of up to RAYS_PER_LAUNCH rays (a whole frame goes
    through ONE set of launches).  Seeded renders are reproducible in either mode; the two modes hand different rows of the same
    uniform stream to a ray, so the frames agree like two jittered renders of one scene (not bit for bit), and the grouped mode
    launches 157x fewer kernels."""
    import dataloader
    import _native as nat
    import nerf_helpers as h
    import numpy as np
    net = make_net(5, "dense")
    H = W = 72
    o, d = dataloader.get_rays(H, W, 0.5 * W / 0.36, h.pose_spherical(-60.0, -30.0, 4.0), device=DEV)
    frames, launches = {}, {}
    try:
        for mode in (None, 1 << 20):
            h.RAYS_PER_LAUNCH = mode
            for rep in range(2):
                torch.manual_seed(123)
                n0 = nat.launches
                frames[(mode, rep)] = h.view_reconstruction(net, o, d, N=1000)          # 5184 rays: 6 chunks, ragged last one
                launches[mode] = nat.launches - n0
    finally:
        h.RAYS_PER_LAUNCH = 1 << 20
    for mode in (None, 1 << 20):
        assert frames[(mode, 0)].dtype.name == "uint8" and frames[(mode, 0)].shape == (H, W, 3)
        assert (frames[(mode, 0)] == frames[(mode, 1)]).all()
    assert launches[None] == 6 * launches[1 << 20] and launches[1 << 20] <= 4
    a, b = frames[(None, 0)].astype(np.float64), frames[(1 << 20, 0)].astype(np.float64)
    psnr = 10 * np.log10(255.0 ** 2 / max(((a - b) ** 2).mean(), 1e-12))
    assert psnr > 30, psnr
