"""Training quality pinned to the reference (nerf_model.py:134-169).

tests/golden/trajectory.npz was written by tests/golden/make_golden.py::gen_trajectory from the UNMODIFIED reference: 200
`training_step` + Adam steps from the synthetic random-init weights on 512-ray batches that are a pure function of the step
(centre-cropped pixels of 8 analytic-scene orbit views, recorded uniforms), the weight set it ends on ("briefly trained"), a
1024-ray step on that set with all 40 gradients, and a 100x100 frame next to the analytic ground truth.

Here the same batches and uniforms go through this repo's bf16 tensor-core training path.  The two runs round differently
(bf16 operands, atomics), and stochastic optimisation amplifies that, so the TRAJECTORIES are compared in bands: the
smoothed loss stays within 5 % of the reference's, the run ends as low as the reference (+10 %), and the densities do not die.
On the reference's trained weights themselves the comparison is tight: forward to the bf16 contract, gradients element-wise.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synthetic
from oracle import nerf_oracle as O
from util import T, rand_triple

pytestmark = pytest.mark.gpu
DEV = "cuda"
VIEWS = 8


def batch(step, rays, images):
    """tests/golden/make_golden.py::trajectory_batch on the device."""
    import dataloader
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    v = step % VIEWS
    c2w = synthetic.orbit_pose(-180.0 + 45.0 * v, -30.0, 4.0)
    xs = (synthetic.uniforms(9000 + 2 * step, (rays,)) * (W // 2)).astype(np.int64) + W // 4
    ys = (synthetic.uniforms(9001 + 2 * step, (rays,)) * (H // 2)).astype(np.int64) + H // 4
    if v not in images:
        images[v] = torch.from_numpy(synthetic.analytic_scene_rgba(c2w.numpy(), H, W, focal)[..., :3].astype(np.float32) / 255.0).to(DEV)
    xs, ys = torch.from_numpy(xs).to(DEV), torch.from_numpy(ys).to(DEV)
    o, d = dataloader.get_rays_at(H, W, focal, c2w, xs, ys)
    return o, d, images[v][ys, xs, :].contiguous()


def trained_net(g):
    import nerf_model
    net = nerf_model.NeRFNetwork()
    net.load_state_dict({k[4:]: T(g[k]) for k in g.files if k.startswith("sd__")})
    return net.to(DEV)


def test_training_trajectory_tracks_the_reference(golden):
    import nerf_model
    g = golden["trajectory"]
    steps, rays = int(g["steps"]), int(g["rays"])
    ref = g["losses"][:, 0]
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(0, "init"))
    net = net.to(DEV)
    opt = net.configure_optimizers()["optimizer"]
    images, losses, nz = {}, [], []
    for step in range(steps):
        o, d, rgb = batch(step, rays, images)
        pred = net.forward(o, d, rand=rand_triple(20000 + 3 * step, rays, device=DEV))
        loss = F.mse_loss(pred["coarse_rgb_rays"], rgb) + F.mse_loss(pred["fine_rgb_rays"], rgb)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.detach())
        nz.append(net.logged["fine_density_non_zeros"])
    losses = torch.stack(losses).cpu().numpy().astype(np.float64)
    nz = torch.stack([x.reshape(()) for x in nz]).cpu().numpy()
    assert abs(losses[0] - ref[0]) < 2e-3 * max(1.0, ref[0])                  # same weights, same batch: the same first loss
    k = 20
    smooth = lambda x: np.convolve(x, np.ones(k) / k, mode="valid")
    a, b = smooth(losses), smooth(ref)
    dev = np.abs(a - b) / b
    print(f"first loss {losses[0]:.5f} (reference {ref[0]:.5f}); last-{k} mean {a[-1]:.5f} (reference {b[-1]:.5f}); "
          f"max smoothed deviation {dev.max():.1%}; fine non-zero densities at the end {nz[-1]:.0f} (reference {g['stats'][-1, 3]:.0f})")
    assert dev.max() < 0.05
    assert a[-1] < 1.10 * b[-1] and a[-1] < 0.5 * losses[0]                    # it learns as fast as the reference does
    assert nz[-1] > 0.25 * g["stats"][-1, 3] and nz[-1] > 0                    # and the density ReLUs are alive


def test_trained_set_forward_and_gradients(golden):
    """The reference's briefly-trained weights, 1024 rays: forward within the bf16 contract (mean 5e-4, max 5e-3, PSNR 60 dB),
    loss to 1e-3 relative, every gradient's norm within 3 % and direction within cosine 0.999 (full tensors for biases / heads,
    16 x 32 corner blocks of the wide matrices element-wise within 5 % of the block's scale)."""
    g = golden["trajectory"]
    net = trained_net(g)
    o, d, target = T(g["g_o"], DEV), T(g["g_d"], DEV), T(g["g_target"], DEV)
    pred = net.forward(o, d, rand=rand_triple(30000, 1024, device=DEV))
    loss = F.mse_loss(pred["coarse_rgb_rays"], target) + F.mse_loss(pred["fine_rgb_rays"], target)
    loss.backward()
    torch.cuda.synchronize()
    for key, ref in (("coarse_rgb_rays", g["g_coarse"]), ("fine_rgb_rays", g["g_fine"])):
        diff = (pred[key].detach().cpu() - T(ref)).abs()
        psnr = 10 * np.log10(1.0 / max(float((diff ** 2).mean()), 1e-20))
        print(f"{key}: max {diff.max():.3e} mean {diff.mean():.3e} PSNR-vs-reference {psnr:.1f} dB")
        assert diff.mean() < 5e-4 and diff.max() < 5e-3 and psnr > 60.0
    ref_loss = float(g["g_loss"])
    assert abs(loss.item() - ref_loss) < 1e-3 * ref_loss, (loss.item(), ref_loss)
    params = dict(net.named_parameters())
    worst_norm, worst_cos, worst_el = 0.0, 1.0, 0.0
    for n, ref_norm in zip([str(x) for x in g["grad_names"]], g["grad_norms"]):
        got = params[n].grad.cpu()
        rel = abs(float(got.norm()) - ref_norm) / max(ref_norm, 1e-12)
        worst_norm = max(worst_norm, rel)
        assert rel < 0.03, f"{n}: |grad| {float(got.norm()):.4e} vs reference {ref_norm:.4e}"
        ref = T(g[f"grad__{n}"])
        blk = got if got.numel() <= 1024 else got[:16, :32]
        cos = F.cosine_similarity(blk.flatten().double(), ref.flatten().double(), dim=0).item()
        el = float((blk - ref).abs().max() / ref.abs().max().clamp(min=1e-20))
        worst_cos, worst_el = min(worst_cos, cos), max(worst_el, el)
        assert cos > 0.999, f"{n}: cosine {cos}"
        assert el < 0.05, f"{n}: element-wise deviation {el:.3f} of the block scale"
    print(f"loss {loss.item():.6f} (reference {ref_loss:.6f}); worst norm deviation {worst_norm:.2%}, worst cosine {worst_cos:.5f}, "
          f"worst element deviation {worst_el:.2%} of block scale")


def test_trained_set_render_and_psnr_delta(golden):
    """100 x 100 frame with the reference's trained weights: PSNR against the reference's own frame >= 60 dB (bf16), and the
    PSNR against the analytic ground truth differs from the reference's by <= 0.05 dB."""
    import dataloader
    import nerf_helpers as h
    g = golden["trajectory"]
    net = trained_net(g)
    o, d = dataloader.get_rays(100, 100, float(g["frame_focal"]), T(g["frame_c2w"]))
    real_rand, state = torch.rand, {"k": 0}

    def fake_rand(shape, device=None, **kw):
        out = T(synthetic.uniforms(31000 + state["k"], tuple(shape)), DEV)
        state["k"] += 1
        return out
    torch.rand = fake_rand
    h.RAYS_PER_LAUNCH = None                # the reference's chunking (one uniform triple per 4096 rays)
    try:
        im = h.view_reconstruction(net, o, d, N=4096)
    finally:
        torch.rand = real_rand
        h.RAYS_PER_LAUNCH = 1 << 20
    vs_ref = O.psnr_uint8(im, g["frame"])
    ours_gt, ref_gt = O.psnr_uint8(im, g["frame_gt"]), O.psnr_uint8(g["frame"], g["frame_gt"])
    print(f"PSNR vs the reference's frame {vs_ref:.2f} dB; vs ground truth {ours_gt:.3f} dB (reference {ref_gt:.3f} dB, delta {ours_gt - ref_gt:+.4f})")
    assert vs_ref > 60.0 and abs(ours_gt - ref_gt) <= 0.05
