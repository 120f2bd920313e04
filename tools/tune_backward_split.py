"""Sweep of the concurrent backward's SM split (training.CONCURRENT_DGRAD_CTAS / CONCURRENT_WGRAD_JOB_CTAS): time of the backward
of both networks for a 4096-ray batch (compositing backward + dgrad || wgrad), against the two kernels in sequence.
usage: python tools/tune_backward_split.py [N]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import torch.nn.functional as F

import nerf_model
import synthetic
import training

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
net = nerf_model.NeRFNetwork()
net.load_state_dict(synthetic.make_state_dict(0, "init"))
net = net.to(dev)
opt = net.configure_optimizers()["optimizer"]
o = torch.randn(N, 3, device=dev) * 0.3
d = F.normalize(torch.randn(N, 3, device=dev), dim=1)
cases = []
for m, S in ((net.coarse_network, 64), (net.fine_network, 192)):
    ts = (2.0 + 4.0 * torch.sort(torch.rand(N, S, 1, device=dev), dim=1).values).contiguous()
    sigma, rgb, acts = training.mlp_forward_train(m, o, d, ts)
    cases.append((m, ts, sigma, rgb, acts))
g = torch.randn(N, 3, device=dev) / N
side = training.side_stream(dev)
main = torch.cuda.current_stream()


def run(split, reps=10):
    def once():
        for m, ts, sigma, rgb, acts in cases:
            training.mlp_backward(m, o, d, ts, sigma, rgb, acts, g, True, split, side)
        if split is not None:
            main.wait_stream(side)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        once()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def shares(total, cost=(25, 17, 17, 18, 17, 18, 17, 12, 7)):
    raw = [c * total / sum(cost) for c in cost]
    n = [max(1, int(r)) for r in raw]
    while sum(n) < total:
        k = max(range(9), key=lambda i: raw[i] - n[i])
        n[k] += 1
    while sum(n) > total:
        k = max(range(9), key=lambda i: n[i] - raw[i] if n[i] > 1 else -1e9)
        n[k] -= 1
    return tuple(n)


print(f"N = {N}: sequential (each kernel on 148 SMs): {run(None):.3f} ms")
costs = {"hbm-bound shares": (25, 17, 17, 18, 17, 18, 17, 12, 7), "flop shares": (14, 16, 16, 16, 16, 16, 16, 10, 3),
         "pe-heavy": (22, 16, 16, 16, 16, 16, 16, 9, 3)}
for gw in (48, 56, 60, 64, 68, 74, 84):
    for name, cost in costs.items():
        sh = shares(gw, cost)
        print(f"dgrad {148 - gw:3d} CTAs | wgrad {gw:3d} CTAs {sh} ({name}): {run((148 - gw, sh)):.3f} ms", flush=True)
