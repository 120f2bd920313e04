"""Three eager 4096-ray training steps (no CUDA graph), the target of the ncu captures of the training kernels:
    ncu --set full --clock-control none --import-source on -k regex:'mlp_tc3|bwd3|wgrad' -s 12 -c 6 -o gpurun_out/train python tools/ncu_train_step.py
(6 tensor-core launches per step: forward coarse / fine, dgrad coarse / fine, wgrad coarse / fine; the third step is captured)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import bench, nerf_model, synthetic, dataloader
from trainer import FlatGradients

dev = torch.device("cuda")
torch.manual_seed(0)
net = nerf_model.NeRFNetwork(); net.load_state_dict(synthetic.make_state_dict(0, "init")); net = net.to(dev)
opt = net.configure_optimizers()["optimizer"]
grads = FlatGradients(net.parameters(), opt)
c2w, focal = bench.frame_setup(800, 800, 3)
image = torch.from_numpy(synthetic.analytic_scene_rgba(c2w.numpy(), 800, 800, focal)[..., :3].copy()).to(dev)
for _ in range(3):
    xs, ys = dataloader.sample_random_coordinates(4096, 800, 800, cropping=True, device=dev)
    o, d = dataloader.get_rays_at(800, 800, focal, c2w, xs, ys)
    rgb = image[ys, xs].float() / 255.0
    grads.zero()
    loss = net.training_step({"origin": o[None], "direc": d[None], "rgb": rgb[None]}, 0)
    loss.backward(); grads.all_reduce_mean(); opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
