"""Kernel-level breakdown of one 4096-ray training step (torch.profiler, CUDA activities)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench, nerf_model, synthetic, dataloader
from trainer import FlatGradients

dev = torch.device("cuda")
net = nerf_model.NeRFNetwork(); net.load_state_dict(synthetic.make_state_dict(0, "init")); net = net.to(dev)
opt = net.configure_optimizers()["optimizer"]
grads = FlatGradients(net.parameters(), opt)
c2w, focal = bench.frame_setup(800, 800, 3)
image = torch.from_numpy(synthetic.analytic_scene_rgba(c2w.numpy(), 800, 800, focal)[..., :3].copy()).to(dev)
def step():
    xs, ys = dataloader.sample_random_coordinates(4096, 800, 800, device=dev)
    o, d = dataloader.get_rays_at(800, 800, focal, c2w, xs, ys)
    rgb = image[ys, xs].float() / 255.0
    grads.zero()
    loss = net.training_step({"origin": o[None], "direc": d[None], "rgb": rgb[None]}, 0)
    loss.backward(); grads.all_reduce_mean(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:22]
tot = sum(e.device_time_total for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA)
print(f"total CUDA time per step: {tot/3/1e3:.3f} ms")
for e in rows:
    if e.device_time_total > 0:
        print(f"{e.device_time_total/3/1e3:8.3f} ms  x{e.count//3:4d}  {e.key[:100]}")
