"""Small end-to-end run for compute-sanitizer (memcheck): every kernel once on tiny shapes - samplers, fp32 MLP, fused
tcgen05 forward (inference + training form), compositing fwd/bwd, dgrad, wgrad."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import torch.nn.functional as F
import dataloader, nerf_model, synthetic

dev = torch.device("cuda")
net = nerf_model.NeRFNetwork()
net.load_state_dict(synthetic.make_state_dict(4, "dense"))
net = net.to(dev)
c2w = synthetic.orbit_pose(30.0, -30.0, 4.0)
o, d = dataloader.get_rays(20, 20, 30.0, c2w)
o, d = o.reshape(-1, 3)[:300].contiguous(), d.reshape(-1, 3)[:300].contiguous()       # 300 rays: ragged tiles everywhere
with torch.no_grad():
    out = net.forward(o, d)
print("inference", out["fine_rgb_rays"].mean().item())
ref = nerf_model.NeRFNetwork(precision="fp32")
ref.load_state_dict(synthetic.make_state_dict(4, "dense"))
ref = ref.to(dev)
with torch.no_grad():
    out32 = ref.forward(o[:40], d[:40])
print("fp32", out32["fine_rgb_rays"].mean().item())
pred = net.forward(o, d)
loss = F.mse_loss(pred["coarse_rgb_rays"], torch.rand(300, 3, device=dev)) + F.mse_loss(pred["fine_rgb_rays"], torch.rand(300, 3, device=dev))
loss.backward()
torch.cuda.synchronize()
print("train", loss.item(), sum(float(p.grad.abs().sum()) for p in net.parameters()))
