"""Worker of tests/test_gpu_dp.py (run under torchrun, one process per GPU, or as a single process).

  entry OUT STEPS SCENE   the product path: `train_nerf.py ... full` (Trainer.fit) for STEPS steps from an UNSEEDED NeRFNetwork();
                          every rank saves its final flat parameter buffer, its initial one and the pixel draws it made
  dp OUT STEPS            data-parallel steps on batches that are a pure function of (rank, step) (hash-based uniforms, no
                          library RNG), through FlatGradients / FlatAdam exactly as Trainer.fit drives them
  single OUT STEPS WORLD  the same steps in ONE process: the batches of all WORLD ranks back to back into the same gradient
                          buffer, grad_scale = 1 / WORLD
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

import dataloader
import nerf_model
import synthetic
from trainer import FlatGradients

RAYS = 1024
H = W = 800


def batch(rank, step, dev):
    """4096... RAYS rays of orbit pose (rank, step) with colours from the analytic scene's formula-free stand-in: a smooth
    function of the pixel, so that no image has to be rendered here."""
    seed = 1000 * rank + 7 * step
    xs = torch.from_numpy((synthetic.uniforms(seed + 1, (RAYS,)) * 400 + 200).astype("int64")).to(dev)
    ys = torch.from_numpy((synthetic.uniforms(seed + 2, (RAYS,)) * 400 + 200).astype("int64")).to(dev)
    c2w = synthetic.orbit_pose(30.0 * rank + 11.0 * step, -30.0, 4.0)
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    o, d = dataloader.get_rays_at(H, W, focal, c2w, xs, ys)
    rgb = torch.stack([(xs % 97).float() / 97, (ys % 89).float() / 89, ((xs + ys) % 53).float() / 53], dim=1)
    rand = tuple(torch.from_numpy(synthetic.uniforms(seed + 10 + k, s)).to(dev) for k, s in enumerate(((RAYS, 64), (RAYS, 1), (RAYS, 128, 1))))
    return o, d, rgb, rand


def loss_of(net, o, d, rgb, rand):
    pred = net.forward(o, d, rand=rand)
    return F.mse_loss(pred['coarse_rgb_rays'], rgb) + F.mse_loss(pred['fine_rgb_rays'], rgb)      # nerf_model.py:159-161


def make_net(dev):
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(0, "init"))
    return net.to(dev)


def main():
    mode, out, steps = sys.argv[1], Path(sys.argv[2]), int(sys.argv[3])
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if mode == "entry":
        import train_nerf
        draws = []
        real = dataloader.sample_random_coordinates

        def spy(*a, **k):
            xs, ys = real(*a, **k)
            draws.append(xs[:8].clone())          # device-side copy: the call may sit inside a CUDA-graph capture (no host sync there)
            return xs, ys
        dataloader.sample_random_coordinates = spy
        first = {}
        real_sync = train_nerf.Trainer.synchronize_replicas

        def sync_spy(self, model, optimizer):
            first["before"] = optimizer.flat_params.clone().cpu()
            real_sync(self, model, optimizer)
            first["after"] = optimizer.flat_params.clone().cpu()
        train_nerf.Trainer.synchronize_replicas = sync_spy
        run = train_nerf.main(["-n", "dp", "--gpu", "-s", str(steps), "-rd", str(out / "exp"), "-r", str(RAYS), "full", "-b", sys.argv[4], "-cr", "1"])
        torch.cuda.synchronize()
        torch.save({"final": run.optimizer.flat_params.cpu(), "before": first["before"], "after": first["after"],
                    "draws": torch.stack(draws).cpu(), "step": run.optimizer._step, "m": run.optimizer.flat_m.cpu()}, out / f"entry_rank{rank}.pt")
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    if mode == "dp":
        dist.init_process_group("nccl", device_id=dev)
        world = dist.get_world_size()
    else:
        world = int(sys.argv[4])
    net = make_net(dev)
    opt = net.configure_optimizers()["optimizer"]
    grads = FlatGradients(net.parameters(), opt)
    n_coarse = sum(p.numel() for p in net.coarse_network.parameters())
    losses = []
    for step in range(steps):
        grads.zero()
        if mode == "dp":
            net.on_coarse_grads_ready = lambda: grads.reduce_async(n_coarse)
            loss = loss_of(net, *batch(rank, step, dev))
            loss.backward()
            grads.all_reduce_mean()
        else:
            opt.grad_scale = 1.0 / world
            for r in range(world):
                loss = loss_of(net, *batch(r, step, dev))
                loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    torch.cuda.synchronize()
    torch.save({"final": opt.flat_params.cpu(), "losses": losses, "grad_scale": opt.grad_scale}, out / f"{mode}_rank{rank}.pt")
    if mode == "dp":
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
