"""A/B on one box: 800x800 frame through render_rays_chunked with and without the side-stream uniform prefetch."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import bench  # noqa: E402
import dataloader  # noqa: E402
import nerf_helpers as h  # noqa: E402
import nerf_model  # noqa: E402
import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
net = nerf_model.NeRFNetwork()
net.load_state_dict(synthetic.make_state_dict(5, "dense"))
net = net.to(dev)
H = W = 800
c2w, focal = bench.frame_setup(H, W, 0)
o, d = dataloader.get_rays(H, W, focal, c2w, device=dev)
o, d = o.reshape(-1, 3), d.reshape(-1, 3)
out = torch.empty_like(o)
for rep in range(3):
    for prefetch in (False, True):
        h.PREFETCH_UNIFORMS = prefetch
        for _ in range(2):
            h.render_rays_chunked(net, o, d, 4096, out=out)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(4):
            h.render_rays_chunked(net, o, d, 4096, out=out)
        t1.record()
        torch.cuda.synchronize()
        print(f"prefetch {prefetch}: {t0.elapsed_time(t1) / 4:.2f} ms / frame", flush=True)
