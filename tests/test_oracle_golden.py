"""Pins oracle/nerf_oracle.py to the reference: golden vectors made by tests/golden/make_golden.py from the
unmodified reference, plus the known-answer values of the reference's own unit tests.  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

import synthetic
from oracle import nerf_oracle as O
from util import T, bits_equal, rand_triple


def test_pose_and_focal(golden):
    g = golden["rays"]
    for k, a in enumerate(g["angles"]):
        assert bits_equal(O.pose_spherical(float(a), -30.0, 4.0), g["poses"][k + 1])
        assert bits_equal(synthetic.orbit_pose(float(a), -30.0, 4.0), g["poses"][k + 1])
    assert O.focal_from_fov(800, 0.6) == float(g["focal_fixture"])
    assert abs(O.focal_from_fov(800, 0.6) - 1293.091257506331) < 5e-8                # dataloader_test.py:39-41 (assertAlmostEqual)
    assert O.focal_from_fov(800, 0.6911112070083618) == float(g["focal800"])


def test_get_rays(golden):
    g = golden["rays"]
    focal800 = float(g["focal800"])
    for k in range(4):
        pose = T(g["poses"][k])
        o, d = O.get_rays(48, 64, 77.25, pose)
        assert bits_equal(d, g[f"d_small_{k}"]) and bits_equal(o.contiguous(), g[f"o_small_{k}"])
        o, d = O.get_rays(800, 800, focal800, pose)
        assert hashlib.sha256((d.contiguous().numpy() + 0.0).tobytes()).hexdigest() == str(g["d800_sha256"][k])
        assert bits_equal(d[799], g[f"d800_row799_{k}"])
    # pixel-list form equals the gather of the full grid
    xs, ys = torch.tensor([0, 5, 63, 17]), torch.tensor([0, 47, 3, 17])
    o, d = O.get_rays(48, 64, 77.25, T(g["poses"][2]), xs, ys)
    assert bits_equal(d, g["d_small_2"][ys.numpy(), xs.numpy()])


def test_coarse_samples(golden):
    g = golden["coarse"]
    o, d = T(g["o"]), T(g["d"])
    pts, ts = O.coarse_samples(o, d, T(synthetic.uniforms(100, (256, 64))), 2.0, 6.0)
    assert bits_equal(pts, g["pts"]) and bits_equal(ts, g["ts"])
    pts, ts = O.coarse_samples(o[:5], d[:5], T(synthetic.uniforms(110, (5, 7))), 0.3, 5.1)
    assert bits_equal(pts, g["pts_odd"]) and bits_equal(ts, g["ts_odd"])


def test_coarse_samples_strata_kat():
    """tests/nerf_helpers_test.py:49-63: o=0, d=1, 2 samples -> one sample strictly inside (2,4) and (4,6)."""
    pts, ts = O.coarse_samples(torch.zeros(1, 3), torch.ones(1, 3), T(synthetic.uniforms(7, (1, 2))) * 0.98 + 0.01)
    lo, hi = torch.tensor([2.0, 4.0]), torch.tensor([4.0, 6.0])
    assert ((lo < ts[0, :, 0]) & (ts[0, :, 0] < hi)).all()
    assert ((lo[:, None] < pts[0]) & (pts[0] < hi[:, None])).all()


def test_compositing(golden):
    g = golden["composite"]
    for S in (64, 192):
        dl = O.deltas(T(g[f"ts_{S}"]))
        assert bits_equal(dl, g[f"deltas_{S}"])
        w = O.weights(T(g[f"sigma_{S}"]), dl)
        assert bits_equal(w, g[f"weights_{S}"])
        assert bits_equal(O.ray_color(w, T(g[f"rgb_{S}"])), g[f"color_{S}"])


def test_compositing_kats(golden):
    # tests/nerf_helpers_test.py:16-21
    dl = torch.full((1, 5, 1), 0.2)
    w = O.weights(torch.tensor([0, 50, 1, 0.3, 1.0]).view(1, 5, 1), dl)
    torch.testing.assert_close(w, torch.tensor([0, 0.9999546001, 8.229611e-6, 2.1646e-6, 6.34545e-6]).view(1, 5, 1))
    assert bits_equal(w, golden["composite"]["kat_weights"])
    # tests/nerf_helpers_test.py:42-47 (int64 ts is promoted)
    dl = O.deltas(torch.arange(2, 6, 1).view(1, -1, 1))
    torch.testing.assert_close(dl, torch.tensor([1, 1, 1, 1e10]).view(1, 4, 1))
    # tests/nerf_helpers_test.py:23-40
    torch.testing.assert_close(O.ray_color(torch.full((1, 256, 1), 1 / 256), torch.ones(1, 256, 3)), torch.ones(1, 3))
    w = torch.zeros(1, 256, 1); w[:, 200] = 1.0
    torch.testing.assert_close(O.ray_color(w, torch.ones(1, 256, 3)), torch.ones(1, 3))


def test_fine_samples(golden):
    g = golden["fine"]
    o, d = T(g["o"]), T(g["d"])
    eps, u = T(synthetic.uniforms(320, (256, 1))), T(synthetic.uniforms(321, (256, 128, 1)))
    pts, fts, idx = O.fine_samples(o, d, T(g["w"]), T(g["c_ts"]), eps, u, return_idx=True)
    assert bits_equal(idx, g["idx"]) and bits_equal(fts, g["f_ts"]) and bits_equal(pts, g["f_pts"])
    assert (g["idx"][0] == 64).all()      # ray 0 has all-zero weights: NaN cdf -> every query lands in the last bin
    eps, u = T(synthetic.uniforms(332, (9, 1))), T(synthetic.uniforms(333, (9, 5, 1)))
    pts, fts = O.fine_samples(o[:9], d[:9], T(g["w_odd"]), T(g["c_ts_odd"]), eps, u)
    assert bits_equal(fts, g["f_ts_odd"]) and bits_equal(pts, g["f_pts_odd"])


def test_positional_encoding(golden):
    g = golden["mlp"]
    x = T(g["pe_x"])
    assert bits_equal(O.positional_encoding(x, 10), g["pe_10"]) and bits_equal(O.positional_encoding(x, 4), g["pe_4"])
    # tests/nerf_model_test.py:41-58: cos block first, then sin
    kat = O.positional_encoding(torch.tensor([[0.0, 0, 0], [1.0, 1, 1]]), 1)
    torch.testing.assert_close(kat, torch.tensor([[1.0, 1, 1, 0, 0, 0], [-1.0, -1, -1, 0, 0, 0]]))
    assert O.positional_encoding(torch.rand(7, 5, 3), 10).shape == (7, 5, 60)      # nerf_model_test.py:60-63


def test_mlp(golden):
    g = golden["mlp"]
    for kind, seed in (("init", 1), ("dense", 2)):
        sd = synthetic.make_state_dict(seed, kind)
        sg, rgb = O.mlp_forward(sd, "fine_network", T(g[f"pts_{kind}"]), T(g[f"dir_{kind}"]))
        assert sg.shape == (8, 16, 1) and rgb.shape == (8, 16, 3)                    # nerf_model_test.py:69-72
        torch.testing.assert_close(sg, T(g[f"sigma_{kind}"]), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(rgb, T(g[f"rgb_{kind}"]), atol=1e-6, rtol=1e-5)


def test_network_forward_and_grads(golden):
    g = golden["network"]
    o, d, target = T(g["o"]), T(g["d"]), T(g["target"])
    o2, d2 = O.get_rays(800, 800, float(g["focal"]), T(g["c2w"]), T(g["xs"]), T(g["ys"]))
    assert bits_equal(o2.contiguous(), o) and bits_equal(d2, d)
    for kind, seed in (("init", 3), ("dense", 4)):
        sd = synthetic.make_state_dict(seed, kind)
        u_c, eps, u_f = rand_triple(500 + seed * 10, 64)
        out = O.network_forward(sd, o, d, u_c, eps, u_f, full=True)
        assert bits_equal(out["ts"], g[f"sorted_ts_{kind}"])
        torch.testing.assert_close(out["coarse_sigma"], T(g[f"c_sigma_{kind}"]), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(out["fine_sigma"], T(g[f"f_sigma_{kind}"]), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(out["fine_rgb_rays"], T(g[f"fine_rgb_rays_{kind}"]), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(out["coarse_rgb_rays"], T(g[f"coarse_rgb_rays_{kind}"]), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(out["stats"], T(g[f"stat_{kind}"]).float(), atol=1e-4, rtol=1e-5)
        loss, grads, _ = O.loss_and_grads(sd, o, d, target, u_c, eps, u_f)
        torch.testing.assert_close(loss.reshape(()), T(g[f"loss_{kind}"]).reshape(()), atol=1e-7, rtol=1e-5)
        assert loss >= 0                                                             # nerf_model_test.py:33-35
        names = [str(n) for n in g[f"grad_names_{kind}"]]
        norms = np.array([float(grads[n].norm()) for n in names])
        np.testing.assert_allclose(norms, g[f"grad_norms_{kind}"], rtol=2e-4, atol=1e-9)
        for n in names:
            key = f"grad_{kind}__{n}"
            if key in g.files:
                torch.testing.assert_close(grads[n], T(g[key]), rtol=2e-4, atol=1e-7 + 1e-5 * float(grads[n].abs().max()))


def test_render_100(golden):
    """BASELINE.json configs[0]: one 100x100 view, chunk 4096, coarse+fine, on CPU."""
    g = golden["render100"]
    sd = synthetic.make_state_dict(5, "dense")
    o, d = O.get_rays(100, 100, float(g["focal"]), T(g["c2w"]))

    def rand_fn(ci, n):
        return rand_triple(600 + 3 * ci, n)
    im = O.view_reconstruction(sd, o.contiguous(), d, rand_fn, N=4096)
    diff = np.abs(im.astype(np.int32) - g["image"].astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01, (diff.max(), (diff > 0).mean())
    assert O.psnr_uint8(im, g["image"]) > 60.0


def test_score_oracle_known_answers():
    """oracle/score_oracle.py restates scikit-image 0.18.3's PSNR / SSIM (score.py:33-37); skimage is not installable here, so
    the restatement is pinned by closed-form cases only."""
    from oracle import score_oracle as S
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(40, 52, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int16) + 1, 0, 255).astype(np.uint8)
    b[a == 255] = 254                                              # |a - b| == 1 everywhere
    assert abs(S.peak_signal_noise_ratio(a, b) - 20 * np.log10(255.0)) < 1e-12
    assert abs(S.structural_similarity(a, a) - 1.0) < 1e-12
    assert abs(S.structural_similarity(a, b) - S.structural_similarity(b, a)) < 1e-12          # symmetric
    # two constant images: variances vanish, S = (2 x y + C1) / (x^2 + y^2 + C1) in every window
    x, y = np.full((16, 16, 3), 100, np.uint8), np.full((16, 16, 3), 140, np.uint8)
    c1 = (0.01 * 255) ** 2
    assert abs(S.structural_similarity(x, y) - (2 * 100 * 140 + c1) / (100 ** 2 + 140 ** 2 + c1)) < 1e-12
    # uncorrelated noise scores far below a lightly perturbed copy
    noisy = np.clip(a.astype(np.int16) + rng.integers(-3, 4, size=a.shape), 0, 255).astype(np.uint8)
    other = rng.integers(0, 256, size=a.shape, dtype=np.uint8)
    assert S.structural_similarity(a, noisy) > 0.98 > 0.2 > S.structural_similarity(a, other)


def test_score_oracle_against_independent_implementations():
    """Two cross-checks that do not share code with oracle/score_oracle.py (still not skimage itself: the oracle stays "parity
    unpinned" in the tier's sense).  PSNR against OpenCV's `cv2.PSNR` (R = 255, all channels pooled - the same definition skimage
    uses for uint8 input).  SSIM against a brute-force evaluation of its definition: every fully-inside 7x7 window, means and SAMPLE
    (ddof = 1) variances / covariance from `np.var` / `np.cov`, S averaged over windows then over channels."""
    cv2 = pytest.importorskip("cv2")
    from oracle import score_oracle as S
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(23, 31, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int16) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    assert abs(S.peak_signal_noise_ratio(a, b) - cv2.PSNR(a, b)) < 1e-9

    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    per_channel = []
    for c in range(3):
        x, y = a[..., c].astype(np.float64), b[..., c].astype(np.float64)
        vals = []
        for i in range(x.shape[0] - 6):
            for j in range(x.shape[1] - 6):
                wx, wy = x[i:i + 7, j:j + 7].ravel(), y[i:i + 7, j:j + 7].ravel()
                ux, uy = wx.mean(), wy.mean()
                vx, vy, vxy = wx.var(ddof=1), wy.var(ddof=1), np.cov(wx, wy, ddof=1)[0, 1]
                vals.append(((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2)))
        per_channel.append(np.mean(vals))
    assert abs(S.structural_similarity(a, b) - np.mean(per_channel)) < 1e-10


def test_oracle_training_path_matches_reference_trajectory(golden):
    """Pins the oracle's loss + autograd path (oracle.training_loss / loss_and_grads) against the UNMODIFIED reference's
    training run (tests/golden/trajectory.npz): the first step's loss on the recorded batch, and the 1024-ray step on the
    briefly-trained weight set - loss, both ray-colour outputs and every gradient the fixture holds."""
    import numpy as np
    import synthetic
    from oracle import nerf_oracle as O
    from util import T, rand_triple
    g = golden["trajectory"]
    rays = int(g["rays"])
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    c2w = synthetic.orbit_pose(-180.0, -30.0, 4.0)
    xs = T((synthetic.uniforms(9000, (rays,)) * (W // 2)).astype(np.int64) + W // 4)
    ys = T((synthetic.uniforms(9001, (rays,)) * (H // 2)).astype(np.int64) + H // 4)
    o, d = O.get_rays(H, W, focal, c2w, xs, ys)
    img = synthetic.analytic_scene_rgba(c2w.numpy(), H, W, focal)[..., :3].astype(np.float32) / 255.0
    target = T(img[ys.numpy(), xs.numpy(), :])
    with torch.no_grad():
        loss0, _ = O.training_loss(synthetic.make_state_dict(0, "init"), o.contiguous(), d, target, *rand_triple(20000, rays))
    assert abs(float(loss0) - g["losses"][0, 0]) < 1e-6
    sd = {k[4:]: T(g[k]) for k in g.files if k.startswith("sd__")}
    loss, grads, out = O.loss_and_grads(sd, T(g["g_o"]), T(g["g_d"]), T(g["g_target"]), *rand_triple(30000, 1024))
    assert abs(float(loss) - float(g["g_loss"])) < 1e-6
    torch.testing.assert_close(out["fine_rgb_rays"], T(g["g_fine"]), atol=1e-6, rtol=0)
    torch.testing.assert_close(out["coarse_rgb_rays"], T(g["g_coarse"]), atol=1e-6, rtol=0)
    for n, ref_norm in zip([str(x) for x in g["grad_names"]], g["grad_norms"]):
        got = grads[n]
        assert abs(float(got.norm()) - ref_norm) <= 1e-4 * ref_norm + 1e-12, n
        blk = got if got.numel() <= 1024 else got[:16, :32]
        torch.testing.assert_close(blk, T(g[f"grad__{n}"]), atol=1e-5 * float(T(g[f"grad__{n}"]).abs().max()) + 1e-12, rtol=1e-3)
