"""GPU parity of the training path: loss and all 40 parameter gradients of one 64-ray step against CPU autograd through
the oracle (golden vectors from the unmodified reference), the compositing backward against torch autograd, and a short
optimisation run that must reduce the loss."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synthetic
from oracle import nerf_oracle as O
from util import T, rand_triple

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_net(seed, kind):
    import nerf_model
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(seed, kind))
    return net.to(DEV)


def test_composite_backward_matches_autograd():
    import training
    g = torch.Generator().manual_seed(3)
    N, S = 37, 192
    ts = (2.0 + 4.0 * torch.sort(torch.rand(N, S, 1, generator=g), dim=1).values)
    sig_pre = torch.randn(N, S, 1, generator=g) * 2 - 0.5
    rgb_pre = torch.randn(N, S, 3, generator=g)
    sig_pre[::4] = -1.0
    gray = torch.randn(N, 3, generator=g)
    a, b = sig_pre.clone().requires_grad_(True), rgb_pre.clone().requires_grad_(True)
    sigma, rgb = torch.relu(a), torch.sigmoid(b)
    col = O.ray_color(O.weights(sigma, O.deltas(ts)), rgb)
    (col * gray).sum().backward()
    dsig, drgb = training.composite_backward(sigma.detach().to(DEV), rgb.detach().to(DEV), ts.to(DEV), gray.to(DEV))
    ref_s, ref_c = a.grad.reshape(-1), b.grad.reshape(-1, 3)
    # compare relative to each ray's gradient scale (all-zero rays must come out exactly zero)
    scale = ref_s.abs().reshape(N, S).max(dim=1, keepdim=True).values.expand(N, S).reshape(-1)
    err = (dsig.cpu() - ref_s).abs()
    assert (err[scale == 0] == 0).all()
    assert (err[scale > 0] / scale[scale > 0]).max() < 1e-4
    torch.testing.assert_close(drgb.cpu(), ref_c, atol=2e-6, rtol=1e-4)


@pytest.mark.parametrize("kind,seed", [("dense", 4)])
def test_training_step_gradients_match_reference(golden, kind, seed):
    g = golden["network"]
    net = make_net(seed, kind)
    o, d, target = T(g["o"], DEV), T(g["d"], DEV), T(g["target"], DEV)
    pred = net.forward(o, d, rand=rand_triple(500 + seed * 10, 64, device=DEV))
    loss = F.mse_loss(pred["coarse_rgb_rays"], target) + F.mse_loss(pred["fine_rgb_rays"], target)   # nerf_model.py:159-161
    loss.backward()
    torch.cuda.synchronize()
    ref_loss = float(g[f"loss_{kind}"])
    assert abs(loss.item() - ref_loss) < 2e-3 * max(1.0, ref_loss), (loss.item(), ref_loss)
    names = [str(n) for n in g[f"grad_names_{kind}"]]
    params = dict(net.named_parameters())
    worst, failures = 0.0, []
    for n, ref_norm in zip(names, g[f"grad_norms_{kind}"]):
        got = params[n].grad
        assert got is not None and got.shape == params[n].shape, n
        rel = abs(float(got.norm()) - ref_norm) / max(ref_norm, 1e-12)
        # 64 rays of the "dense" synthetic weight set: every ray is opaque, so the density head's gradient is the small
        # difference of large terms (|grad| 1.7e-5 next to 1.7e-2 for rgb_fn.0.weight) and moves by several per cent with the bf16
        # rounding of its operands; everything else is held to 3 %.  (The 1024-ray step on the briefly-trained weights,
        # tests/test_gpu_trajectory.py, holds ALL 40 norms to 3 % and measures 1.2 %.)
        tol = 0.10 if "density_fn" in n else 0.03
        worst = max(worst, rel)
        if rel >= tol:
            failures.append(f"{n}: |grad| {float(got.norm()):.4e} vs reference {ref_norm:.4e} ({rel:.1%})")
        key = f"grad_{kind}__{n}"
        if key in g.files:                      # full tensors for biases and the two small heads
            ref = T(g[key])
            cos = F.cosine_similarity(got.cpu().flatten(), ref.flatten(), dim=0).item()
            # 64 rays only: the first layer's bias gradient sums 64 x 64 / 64 x 192 bf16-rounded dz rows and measures 0.9979 here;
            # the 1024-ray step on the trained weight set holds every gradient to cosine 0.999
            if cos <= 0.997:
                failures.append(f"{n}: cosine {cos:.5f}")
        else:
            ref = T(g[f"gradhead_{kind}__{n}"])
            torch.testing.assert_close(got.cpu()[:4, :8], ref, rtol=0.1, atol=0.15 * float(ref.abs().max()) + 1e-9)   # bf16 chain: element-wise within 15 % of the block scale
    assert not failures, "\n".join(failures)
    print(f"loss {loss.item():.6f} (reference {ref_loss:.6f}); worst gradient-norm deviation {worst:.3%}")


def test_gradient_run_to_run_spread_is_atomics_noise(golden):
    """wgrad adds its per-CTA partial sums with fp32 atomics, so gradients are not bit-reproducible: bound the spread.  Five
    backward passes over the same batch: every parameter's gradient stays within 1e-5 (relative 2-norm) of the first run's."""
    g = golden["network"]
    net = make_net(4, "dense")
    o, d, target = T(g["o"], DEV), T(g["d"], DEV), T(g["target"], DEV)
    runs = []
    for _ in range(5):
        net.zero_grad(set_to_none=True)
        pred = net.forward(o, d, rand=rand_triple(540, 64, device=DEV))
        (F.mse_loss(pred["coarse_rgb_rays"], target) + F.mse_loss(pred["fine_rgb_rays"], target)).backward()
        runs.append([p.grad.double().clone() for p in net.parameters()])
    worst = 0.0
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            worst = max(worst, float((a - b).norm() / a.norm().clamp(min=1e-30)))
    print(f"worst relative gradient difference between identical runs: {worst:.2e}")
    assert worst < 1e-5


def test_deterministic_wgrad_is_bit_reproducible_and_agrees_with_split_k(golden, monkeypatch):
    """training.DETERMINISTIC_WGRAD: one CTA per wgrad job, one add per gradient element.  Three backward passes over the same
    batch give bit-identical gradients, and they agree with the default split-K form to 1e-3 (relative 2-norm per parameter: the
    single CTA's fp32 accumulators in TMEM run over every tile of the batch, a longer summation chain than a split-K share's)."""
    import training
    g = golden["network"]
    net = make_net(4, "dense")
    o, d, target = T(g["o"], DEV), T(g["d"], DEV), T(g["target"], DEV)

    def grads():
        net.zero_grad(set_to_none=True)
        pred = net.forward(o, d, rand=rand_triple(540, 64, device=DEV))
        (F.mse_loss(pred["coarse_rgb_rays"], target) + F.mse_loss(pred["fine_rgb_rays"], target)).backward()
        return [p.grad.clone() for p in net.parameters()]

    split_k = grads()
    monkeypatch.setattr(training, "DETERMINISTIC_WGRAD", True)
    runs = [grads() for _ in range(3)]
    for other in runs[1:]:
        for (name, _), a, b in zip(net.named_parameters(), runs[0], other):
            assert torch.equal(a, b), name
    worst = 0.0
    for (name, _), a, b in zip(net.named_parameters(), runs[0], split_k):
        rel = float((a.double() - b.double()).norm() / a.double().norm().clamp(min=1e-30))
        worst = max(worst, rel)
        assert rel < 1e-3, (name, rel)
    print(f"deterministic vs split-K wgrad: worst relative gradient difference {worst:.2e}")


def test_forward_only_networks_fail_backward_with_the_reason():
    """precision='fp32' / other encoding sizes run the forward-only exact kernel: forward works with gradients enabled, the
    backward raises a RuntimeError that says why; a training run refuses them up front (train_nerf.py -p / -d)."""
    import nerf_model
    import trainer
    for kw in (dict(precision="fp32"), dict(position_dim=6, direction_dim=2)):
        net = nerf_model.NeRFNetwork(**kw).to(DEV)
        o = torch.zeros(8, 3, device=DEV)
        d = F.normalize(torch.ones(8, 3, device=DEV), dim=1)
        out = net.forward(o, d)
        assert out["fine_rgb_rays"].shape == (8, 3) and torch.isfinite(out["fine_rgb_rays"]).all()
        with pytest.raises(RuntimeError, match="differentiable path exists"):
            out["fine_rgb_rays"].sum().backward()
        with pytest.raises(RuntimeError, match="differentiable path exists"):
            trainer.Trainer(max_steps=1, save_checkpoints=False).fit(net, train_dataloaders=[])


def test_no_gradient_from_fine_loss_into_coarse_network(golden):
    g = golden["network"]
    net = make_net(4, "dense")
    pred = net.forward(T(g["o"], DEV), T(g["d"], DEV), rand=rand_triple(540, 64, device=DEV))
    F.mse_loss(pred["fine_rgb_rays"], T(g["target"], DEV)).backward()
    assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in net.coarse_network.parameters())
    assert any(float(p.grad.abs().max()) > 0 for p in net.fine_network.parameters())


def test_short_optimisation_reduces_loss():
    """training_step + Adam from configure_optimizers (nerf_model.py:134-169) on a fixed synthetic batch."""
    import dataloader
    torch.manual_seed(0)
    net = make_net(7, "dense")
    opt = net.configure_optimizers()["optimizer"]
    c2w = synthetic.orbit_pose(40.0, -30.0, 4.0)
    focal = O.focal_from_fov(800, 0.6911112070083618)
    xs = torch.randint(200, 600, (1024,), device=DEV)
    ys = torch.randint(200, 600, (1024,), device=DEV)
    o, d = dataloader.get_rays_at(800, 800, focal, c2w, xs, ys)
    img = torch.from_numpy(synthetic.analytic_scene_rgba(c2w.numpy(), 800, 800, focal)[..., :3].copy()).to(DEV)
    rgb = img[ys, xs].float() / 255.0
    losses = []
    for step in range(30):
        batch = {"origin": o[None], "direc": d[None], "rgb": rgb[None]}
        loss = net.training_step(batch, step)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("loss trajectory:", " ".join(f"{l:.4f}" for l in losses[::5]), f"-> {losses[-1]:.4f}")
    assert np.isfinite(losses).all() and losses[-1] < 0.7 * losses[0]


def test_dgrad_kernel_matches_library_chain():
    """Hand-written tcgen05 dgrad chain vs the same chain as cuBLAS GEMMs + elementwise ops, several tiles per CTA and a
    ragged last tile: all 20 gradients of one network."""
    import library_backward
    import training
    torch.manual_seed(1)
    net = make_net(4, "dense")
    N, S = 700, 67                                        # 46 900 samples = 366.4 tiles
    o = torch.randn(N, 3, device=DEV) * 0.3
    d = F.normalize(torch.randn(N, 3, device=DEV), dim=1) * 1.05
    ts = (2.0 + 4.0 * torch.sort(torch.rand(N, S, 1, device=DEV), dim=1).values).contiguous()
    model = net.fine_network
    sigma, rgb, acts = training.mlp_forward_train(model, o, d, ts)
    g_ray = torch.randn(N, 3, device=DEV) / N
    got = training.mlp_backward(model, o, d, ts, sigma, rgb, acts, g_ray)                   # tcgen05 dgrad + tcgen05 wgrad
    mid = library_backward.mlp_backward_library_wgrad(model, o, d, ts, sigma, rgb, acts, g_ray)     # tcgen05 dgrad + cuBLAS wgrad
    ref = library_backward.mlp_backward_reference(model, o, d, ts, sigma, rgb, acts, g_ray)         # everything as library ops
    torch.cuda.synchronize()
    names = [n for n, _ in model.named_parameters()]
    for name, a, m, b in zip(names, got, mid, ref):
        a, m, b = a.float().flatten(), m.float().flatten(), b.float().flatten()
        for tag, x, y, tol in (("wgrad kernel vs library wgrad", a, m, 0.01), ("kernels vs library chain", a, b, 0.03)):
            cos = F.cosine_similarity(x, y, dim=0).item()
            rel = ((x - y).norm() / y.norm().clamp(min=1e-20)).item()
            assert cos > 0.999 and rel < tol, f"{name} ({tag}): cosine {cos:.5f} rel {rel:.4f}"


def test_graphed_train_step():
    """training.GraphedTrainStep: Trainer.fit's step body captured as one CUDA graph.  Replays train (the loss falls like the
    eager step's), the optimiser's host and device step counts stay together, a learning-rate change by the scheduler reaches the
    captured Adam kernel through device memory, and the packed bf16 weight images follow the parameters (a render with the
    trained network equals a render with a fresh network loaded from its state_dict, bit for bit)."""
    import nerf_model
    import training
    from trainer import FlatGradients
    H = W = 800
    focal = O.focal_from_fov(W, 0.6911112070083618)
    poses = torch.stack([synthetic.orbit_pose(a, -30.0, 4.0) for a in (40.0, 130.0)]).to(DEV).contiguous()
    images = torch.stack([torch.from_numpy(synthetic.analytic_scene_rgba(p.cpu().numpy(), H, W, focal)[..., :3].copy()) for p in poses]).to(DEV).contiguous()
    results = {}
    for mode in ("graph", "eager"):
        torch.manual_seed(7)
        net = make_net(0, "init")
        opt = net.configure_optimizers()["optimizer"]
        grads = FlatGradients(net.parameters(), opt)
        stepper = training.GraphedTrainStep(net, opt, grads, images, poses, focal, 1024, cropping=True)
        losses = [float(stepper.first_loss)]
        for k in range(40):
            loss = stepper.step(k % 2) if mode == "graph" else stepper.eager_step(k % 2)
            losses.append(float(loss))
        torch.cuda.synchronize()
        assert opt._step == 43 and int(opt.dev_step) == 43, (opt._step, int(opt.dev_step))
        results[mode] = (losses, net, opt, stepper)
    lg, le = results["graph"][0], results["eager"][0]
    print("graph:", " ".join(f"{l:.4f}" for l in lg[::8]), "| eager:", " ".join(f"{l:.4f}" for l in le[::8]))
    assert np.isfinite(lg).all() and np.mean(lg[-5:]) < 0.6 * lg[0] and np.mean(le[-5:]) < 0.6 * le[0]
    assert abs(np.mean(lg[-10:]) - np.mean(le[-10:])) < 0.35 * np.mean(le[-10:])
    # the scheduler's learning rate reaches the captured kernel: lr = 0 freezes the parameters
    _, net, opt, stepper = results["graph"]
    before = opt.flat_params.clone()
    opt.param_groups[0]["lr"] = 0.0
    stepper.step(0)
    assert torch.equal(opt.flat_params, before)
    opt.param_groups[0]["lr"] = 5e-4
    stepper.step(1)
    assert not torch.equal(opt.flat_params, before)
    # packed weight images in step with the parameters
    fresh = nerf_model.NeRFNetwork()
    fresh.load_state_dict({k: v.detach().clone() for k, v in net.state_dict().items()})
    fresh = fresh.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    o = torch.randn(300, 3, device=DEV, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 4.0], device=DEV)
    d = F.normalize(-o + 0.3 * torch.randn(300, 3, device=DEV, generator=g), dim=1)
    rand = rand_triple(77, 300, device=DEV)
    with torch.no_grad():
        a = net.forward(o, d, rand=rand)["fine_rgb_rays"]
        b = fresh.forward(o, d, rand=rand)["fine_rgb_rays"]
    assert torch.equal(a, b)


@pytest.mark.parametrize("N,S", [(700, 67), (4096, 64), (1024, 192), (1, 64), (37, 1), (5, 1024), (301, 192)])
def test_dgrad_with_compositing_backward_inside_is_bit_identical(N, S):
    """nerf_mlp_backward_tc_fused (the compositing backward in the dgrad kernel's producer warps) against nerf_composite_backward +
    nerf_mlp_backward_tc: the same per-ray routine, so the dz tensor - every layer's pre-activation gradient and the heads block -
    must be BIT-identical, for rays that straddle tile pairs (S = 67, 192), many rays per pair (S = 1, 64) and rays longer than a
    pair (S = 1024)."""
    import _native as nat
    import training
    torch.manual_seed(3)
    net = make_net(4, "dense")
    o = torch.randn(N, 3, device=DEV) * 0.3
    d = F.normalize(torch.randn(N, 3, device=DEV), dim=1) * 1.05
    ts = (2.0 + 4.0 * torch.sort(torch.rand(N, S, 1, device=DEV), dim=1).values).contiguous()
    model = net.fine_network
    sigma, rgb, (acts, masks) = training.mlp_forward_train(model, o, d, ts)
    g_ray = torch.randn(N, 3, device=DEV) / N
    M = N * S
    lib = nat.lib()
    dsig, drgb = training.composite_backward(sigma, rgb, ts, g_ray)
    a = torch.zeros((training.padded_rows(M) * training.DZ,), device=DEV, dtype=torch.bfloat16)
    b = torch.zeros_like(a)
    packed_t = model.packed_weights_t()
    nat.check(lib.nerf_mlp_backward_tc(nat.ptr(packed_t), nat.ptr(masks), nat.ptr(dsig), nat.ptr(drgb), N, S, nat.ptr(a), nat.stream()), "dgrad")
    nat.check(lib.nerf_mlp_backward_tc_fused(nat.ptr(packed_t), nat.ptr(masks), nat.ptr(sigma), nat.ptr(rgb), nat.ptr(ts), nat.ptr(g_ray), N, S,
                                             nat.ptr(b), nat.stream()), "dgrad fused")
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    assert float(a.float().abs().max()) > 0


def test_flat_adam_matches_torch_adam():
    """optim.FlatAdam (one hand-written kernel over flat buffers, csrc/adam.cu) against torch.optim.Adam on the same
    gradients for 25 steps with a decaying learning rate, then a state_dict round trip into a fresh optimiser (the
    PL-format checkpoints store optimizer.state_dict())."""
    import copy
    import optim
    torch.manual_seed(3)
    net_a = make_net(4, "dense")
    net_b = copy.deepcopy(net_a)
    opt_a = optim.FlatAdam(net_a.parameters(), lr=5e-4)
    opt_b = torch.optim.Adam(net_b.parameters(), lr=5e-4)
    sch_a = torch.optim.lr_scheduler.ExponentialLR(opt_a, gamma=0.9)
    sch_b = torch.optim.lr_scheduler.ExponentialLR(opt_b, gamma=0.9)
    gen = torch.Generator(device=DEV).manual_seed(11)

    def run(steps, oa, ob, sa, sb):
        for _ in range(steps):
            oa.zero_grad()
            for pa, pb in zip(net_a.parameters(), net_b.parameters()):
                g = torch.randn(pa.shape, device=DEV, generator=gen) * (10.0 ** float(torch.randint(-6, 1, (1,)).item()))
                pa.grad.copy_(g)
                pb.grad = g.clone()
            oa.step(); ob.step(); sa.step(); sb.step()

    run(25, opt_a, opt_b, sch_a, sch_b)
    for (name, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
        torch.testing.assert_close(pa.detach(), pb.detach(), rtol=2e-6, atol=1e-7, msg=lambda m: f"{name}: {m}")
    sd = opt_a.state_dict()
    assert set(sd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"} and sd["param_groups"][0]["betas"] == (0.9, 0.999)
    opt_c = optim.FlatAdam(net_a.parameters(), lr=5e-4)
    opt_c.load_state_dict(copy.deepcopy(sd))
    sch_c = torch.optim.lr_scheduler.ExponentialLR(opt_c, gamma=0.9)
    sch_c.load_state_dict(sch_a.state_dict())
    run(5, opt_c, opt_b, sch_c, sch_b)
    for (name, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
        torch.testing.assert_close(pa.detach(), pb.detach(), rtol=5e-6, atol=1e-7, msg=lambda m: f"{name} after resume: {m}")


def test_graph_safe_adam_matches_the_host_scalar_form():
    """nerf_adam_step_dev (lr / betas / eps / grad_scale / step count in device memory, bias corrections evaluated on the device:
    what the replayed CUDA graph runs) against nerf_adam_step (host scalars) and torch.optim.Adam: 30 steps with a decaying
    learning rate and grad_scale = 1/4 (data parallel), parameters equal to 2e-6 relative."""
    import copy
    import optim
    torch.manual_seed(5)
    net_a = make_net(4, "dense")
    net_b, net_c = copy.deepcopy(net_a), copy.deepcopy(net_a)
    opt_a = optim.FlatAdam(net_a.parameters(), lr=5e-4)
    opt_b = optim.FlatAdam(net_b.parameters(), lr=5e-4)
    opt_c = torch.optim.Adam(net_c.parameters(), lr=5e-4)
    opt_a.graph_safe = True
    opt_a.grad_scale = opt_b.grad_scale = 0.25
    opt_a.sync_device_state(); opt_a.set_device_step()
    gen = torch.Generator(device=DEV).manual_seed(12)
    for step in range(30):
        lr = 5e-4 * 0.93 ** step
        for o_ in (opt_a, opt_b, opt_c):
            o_.param_groups[0]["lr"] = lr
        for pa, pb, pc in zip(net_a.parameters(), net_b.parameters(), net_c.parameters()):
            g = torch.randn(pa.shape, device=DEV, generator=gen) * (10.0 ** float(torch.randint(-6, 1, (1,)).item()))
            pa.grad.copy_(g); pb.grad.copy_(g); pc.grad = g * 0.25
        opt_a.step(); opt_b.step(); opt_c.step()
    assert opt_a._step == 30 and int(opt_a.dev_step) == 30
    for (name, pa), pb, pc in zip(net_a.named_parameters(), net_b.parameters(), net_c.parameters()):
        torch.testing.assert_close(pa.detach(), pb.detach(), rtol=2e-6, atol=1e-7, msg=lambda m: f"{name} (device vs host scalars): {m}")
        torch.testing.assert_close(pa.detach(), pc.detach(), rtol=5e-6, atol=1e-7, msg=lambda m: f"{name} (vs torch.optim.Adam): {m}")


def test_batch_rays_matches_the_dataset_item_path(tmp_path):
    """nerf_batch_rays (image index read from DEVICE memory; the batch producer inside the replayed step) against the per-item path
    of SyntheticDataset.__getitem__ (nerf_raygen for the pixel list + colour gather): rays bit-identical, colours = fl32(u8 / 255)
    with the division in double precision as numpy's `imread(...) / 255` upstream (dataloader.py:148)."""
    import dataloader
    synthetic.write_blender_scene(tmp_path, n_train=3, n_val=1, n_test=1)
    ds = dataloader.SyntheticDataset(tmp_path, "train", 777, cropping=False)
    images, poses = ds.stacked()
    assert images.shape == (3, 800, 800, 3) and images.dtype == torch.uint8 and poses.shape == (3, 4, 4)
    g = torch.Generator(device=DEV).manual_seed(8)
    xs = torch.randint(0, 800, (777,), device=DEV, generator=g)
    ys = torch.randint(0, 800, (777,), device=DEV, generator=g)
    idx = torch.zeros((), device=DEV, dtype=torch.int64)
    for k in (2, 0, 1):
        idx.fill_(k)
        o, d, rgb = dataloader.batch_rays(poses, idx, images, ds.focal, xs, ys)
        o_ref, d_ref = dataloader.get_rays_at(800, 800, ds.focal, torch.tensor(ds.frames[k]["transform_matrix"], dtype=torch.float32), xs, ys)
        assert torch.equal(o, o_ref) and torch.equal(d, d_ref)
        want = torch.from_numpy((ds.image_u8(k)[ys, xs].cpu().numpy().astype(np.float64) / 255).astype(np.float32))
        assert torch.equal(rgb.cpu(), want)


def test_gradients_are_linear_in_the_upstream_gradient():
    """Full batch size (4096 rays, both sample counts): doubling dL/d(ray colour) doubles every parameter gradient - the compositing
    backward, the dgrad chain (bf16 dz: a factor 2 is exact) and wgrad are linear for fixed ReLU masks; what is left is the order of
    the fp32 atomics (1e-6)."""
    import training
    torch.manual_seed(6)
    net = make_net(0, "init")
    N = 4096
    o = torch.randn(N, 3, device=DEV) * 0.2 + torch.tensor([0.0, 0.0, 4.0], device=DEV)
    d = F.normalize(-o + 0.4 * torch.randn(N, 3, device=DEV), dim=1)
    for model, S in ((net.coarse_network, 64), (net.fine_network, 192)):
        ts = (2.0 + 4.0 * torch.sort(torch.rand(N, S, 1, device=DEV), dim=1).values).contiguous()
        sigma, rgb, acts = training.mlp_forward_train(model, o, d, ts)
        g_ray = torch.randn(N, 3, device=DEV) / N
        g1 = training.mlp_backward(model, o, d, ts, sigma, rgb, acts, g_ray)
        g2 = training.mlp_backward(model, o, d, ts, sigma, rgb, acts, 2.0 * g_ray)
        torch.cuda.synchronize()
        for name, a, b in zip([n for n, _ in model.named_parameters()], g1, g2):
            rel = ((b.double() - 2.0 * a.double()).norm() / (2.0 * a.double().norm()).clamp(min=1e-30)).item()
            assert rel < 1e-5, f"{name} (S = {S}): relative deviation from linearity {rel:.2e}"


def test_repack_all_matches_individual_packs():
    """nerf_pack_weights_all (the optimiser's post-step hook: four images, one launch) writes exactly what the per-image packers
    write, and the cached images are picked up by the next forward / backward."""
    import nerf_model
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(4, "dense"))
    net = net.to(DEV)
    nets = (net.coarse_network, net.fine_network)
    before = [n.packed_weights().clone() for n in nets]
    with torch.no_grad():
        for p in net.parameters():
            p.data.mul_(1.03)                       # in place through .data: no version bump, like the flat Adam kernel
    net.repack_all()
    got = [t.clone() for n in nets for t in (n._packed, n._packed_t)]
    assert all(n.packed_weights() is n._packed for n in nets)               # keys are current: no re-pack on use
    net.invalidate_packed_weights()
    ref = [t.clone() for n in nets for t in (n.packed_weights(), n.packed_weights_t())]
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(got, ref))
    assert not torch.equal(got[0], before[0]) and not torch.equal(got[2], before[1])
