"""Does interleaving ray chunks over two CUDA streams hide the small kernels / launch gaps between the persistent MLP kernels?
Renders one 800x800 frame (157 chunks of 4096 rays) on 1, 2 and 3 streams and prints ms per frame."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import bench  # noqa: E402
import dataloader  # noqa: E402
import nerf_model  # noqa: E402
import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
net = nerf_model.NeRFNetwork()
net.load_state_dict(synthetic.make_state_dict(5, "dense"))
net = net.to(dev)
H = W = 800
n = H * W
c2w, focal = bench.frame_setup(H, W, 0)
o, d = dataloader.get_rays(H, W, focal, c2w, device=dev)
o, d = o.reshape(n, 3), d.reshape(n, 3)
out = torch.empty((n, 3), device=dev)
CH = 4096


def render(ns):
    main = torch.cuda.current_stream()
    if ns == 1:
        with torch.no_grad():
            for i in range(0, n, CH):
                out[i:i + CH] = net.forward(o[i:i + CH], d[i:i + CH])["fine_rgb_rays"]
        return
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for s in streams:
        s.wait_stream(main)
    with torch.no_grad():
        for k, i in enumerate(range(0, n, CH)):
            with torch.cuda.stream(streams[k % ns]):
                out[i:i + CH] = net.forward(o[i:i + CH], d[i:i + CH])["fine_rgb_rays"]
    for s in streams:
        main.wait_stream(s)


for ns in (1, 2, 3, 1, 2):
    for _ in range(2):
        render(ns)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    for _ in range(3):
        render(ns)
    t1.record()
    torch.cuda.synchronize()
    print(f"streams {ns}: {t0.elapsed_time(t1) / 3:.2f} ms / frame (wall {(time.perf_counter() - w0) / 3 * 1e3:.2f})", flush=True)
