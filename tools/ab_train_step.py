"""Same-box A/B of the replayed training step (training.GraphedTrainStep, 4096 rays) for module-level switches of training.py:
    python tools/ab_train_step.py FUSE_COMPOSITE_BACKWARD=1 FUSE_COMPOSITE_BACKWARD=0
Each variant: fresh network, capture, 60 replays timed with CUDA events; variants are interleaved twice."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
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


def run(setting):
    name, val = setting.split("=")
    setattr(training, name, type(getattr(training, name))(int(val)))
    torch.manual_seed(1)
    net = nerf_model.NeRFNetwork(); net.load_state_dict(synthetic.make_state_dict(0, "init")); net = net.to(dev)
    opt = net.configure_optimizers()["optimizer"]
    grads = FlatGradients(net.parameters(), opt)
    st = training.GraphedTrainStep(net, opt, grads, images, poses, focal, 4096, cropping=True)
    for k in range(10):
        st.step(k % 2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(60):
        loss = st.step(k % 2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 60
    print(f"{setting:40s} {ms:.4f} ms per step   loss {float(loss):.5f}", flush=True)
    st.close()


for rep in range(2):
    for setting in sys.argv[1:]:
        run(setting)
