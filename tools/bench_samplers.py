"""Achieved HBM GB/s of the sampling / compositing kernels at BASELINE sizes vs the measured copy bandwidth.
Algorithmic bytes per launch are the figures of DESIGN.md section 4.2 (inputs read once + outputs written once)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import dataloader, nerf_helpers as h, synthetic, training

dev = torch.device("cuda")
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
N, C, F = 1 << 18, 64, 128            # 262 144 rays per launch so every kernel moves far more than the 126 MB L2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ms = []
    for i in range(iters):
        flush.fill_(i); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return min(ms)

o = torch.randn(N, 3, device=dev) * 0.2
d = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=1) * 1.05
u_c = torch.rand(N, C, device=dev)
_, ts = h.generate_coarse_samples(o, d, C, rand=u_c)
sigma = torch.relu(torch.randn(N, C, 1, device=dev) * 3 - 2)
rgb = torch.rand(N, C, 3, device=dev)
w = h.composite(sigma, rgb, ts)["weights"]
eps, u_f = torch.rand(N, 1, device=dev), torch.rand(N, F, 1, device=dev)
_, fts = h.inverse_transform_sampling(o, d, w, ts, F, rand=(eps, u_f))
S = C + F
ts_all = torch.sort(torch.cat([fts, ts], 1), 1).values.contiguous()
sig2, rgb2 = torch.relu(torch.randn(N, S, 1, device=dev) * 3 - 2), torch.rand(N, S, 3, device=dev)
g = torch.randn(N, 3, device=dev)
c2w = synthetic.orbit_pose(10.0, -30.0, 4.0)
rows = [
    ("raygen_kernel (800x800 grid x 40; host-call bound: ~26 us per get_rays call)", lambda: [dataloader.get_rays(800, 800, 1111.1, c2w) for _ in range(40)], 40 * 640000 * 24),
    ("coarse_sample_kernel (ts only)", lambda: training and __import__("nerf_model"), 0),
]
import nerf_model
net = nerf_model.NeRFNetwork().to(dev)
rows[1] = ("coarse_sample_kernel (ts only)", lambda: net._coarse_ts(o, d, u_c), N * (24 + 4 * C + 4 * C))
rows += [
    ("composite_kernel<1> S=64 (+weights)", lambda: h.composite(sigma, rgb, ts), N * (20 * C + 4 * C + 20)),
    ("composite_kernel<1> S=192", lambda: h.composite(sig2, rgb2, ts_all, want_weights=False), N * (20 * S + 20)),
    ("fine_sample_kernel (ts only would be less; writes pts+ts)", lambda: h.inverse_transform_sampling(o, d, w, ts, F, rand=(eps, u_f)),
     N * (24 + 4 * (2 * C + 1 + F) + 16 * F)),
    ("fine_sample_merge_kernel (K3 + K4 in one launch, sorted ts only)", lambda: h.fine_depths_sorted(w, ts, F, rand=(eps, u_f)),
     N * (4 * (2 * C + 1 + F) + 4 * S)),
    ("merge_sort_kernel (ts only)", lambda: h.merge_samples(o, d, fts, ts, want_points=False), N * (4 * S + 4 * S)),
    ("composite_backward_kernel S=192", lambda: training.composite_backward(sig2, rgb2, ts_all, g), N * (20 * S + 12 + 16 * S)),
]
print(f"HBM copy peak (measured): {peak:.0f} GB/s; N = {N} rays per launch")
for name, fn, nbytes in rows:
    ms = timed(fn)
    gbs = nbytes / ms / 1e6
    print(f"{name:80s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f} % of peak")
