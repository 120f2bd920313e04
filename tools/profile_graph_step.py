"""Timeline of ONE replayed training step (training.GraphedTrainStep, 4096 rays): every kernel with its duration and the idle gap
before it (torch.profiler / CUPTI on a graph replay).  usage: python tools/profile_graph_step.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench, nerf_model, synthetic, training
from trainer import FlatGradients

dev = torch.device("cuda")
H = W = 800
images, poses = [], []
for j in range(2):
    c2w, focal = bench.frame_setup(H, W, 3 + 20 * j)
    poses.append(c2w.to(torch.float32))
    images.append(torch.from_numpy(synthetic.analytic_scene_rgba(c2w.numpy(), H, W, focal)[..., :3].copy()))
images, poses = torch.stack(images).to(dev).contiguous(), torch.stack(poses).to(dev).contiguous()
net = nerf_model.NeRFNetwork(); net.load_state_dict(synthetic.make_state_dict(0, "init")); net = net.to(dev)
opt = net.configure_optimizers()["optimizer"]
grads = FlatGradients(net.parameters(), opt)
st = training.GraphedTrainStep(net, opt, grads, images, poses, focal, 4096, cropping=True)
for k in range(10):
    st.step(k % 2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for k in range(3):
        st.step(k % 2)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
# split into replays: the fill_ of the image index precedes each replay
starts = [i for i, e in enumerate(ev) if "FillFunctor<long" in e.name or "fill" in e.name.lower() and e.time_range.elapsed_us() < 5]
last = ev[len(ev) * 2 // 3:]                       # roughly the third replay
t0 = last[0].time_range.start
busy, prev_end, gaps = 0.0, None, 0.0
print(f"{'start us':>9s} {'dur us':>8s} {'gap us':>7s}  kernel")
for e in last:
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
    gaps += max(gap, 0.0)
    busy += e.time_range.elapsed_us()
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.elapsed_us():8.1f} {gap:7.1f}  {e.name[:90]}")
    prev_end = e.time_range.end
print(f"kernels {len(last)}  busy {busy / 1e3:.3f} ms  gaps {gaps / 1e3:.3f} ms  span {(last[-1].time_range.end - t0) / 1e3:.3f} ms")
st.close()
