"""Per-CTA elapsed clocks of wgrad_tc_kernel by job group (load balance of the fixed CTA split)."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import numpy as np
import torch
import _native as nat, nerf_model, synthetic, training
dev = torch.device("cuda")
net = nerf_model.NeRFNetwork(); net.load_state_dict(synthetic.make_state_dict(4, "dense")); net = net.to(dev)
N, S = 4096, 192
m = net.fine_network
o = torch.randn(N, 3, device=dev) * 0.3
d = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=1)
ts = (2 + 4 * torch.sort(torch.rand(N, S, 1, device=dev), dim=1).values).contiguous()
sigma, rgb, acts = training.mlp_forward_train(m, o, d, ts)
g = torch.randn(N, 3, device=dev) / N
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nat.kernel_events = []
    grads = training.mlp_backward(m, o, d, ts, sigma, rgb, acts, g)
    torch.cuda.synchronize()
    for name, units, t0, t1 in nat.kernel_events:
        print(f"{name}: {t0.elapsed_time(t1):.3f} ms")
fn = nat.debug_lib().nerf_debug_wgrad_cycles
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_longlong * 320)()
assert fn(buf) == 0
c = np.array(buf[:]).reshape(160, 2)[:148]
ctas = [25, 17, 17, 18, 17, 18, 17, 12, 7]          # wg::c_jobs[].ctas
names = ["PE(x) parts of mlp.0/ff.0", "mlp.2", "mlp.4", "mlp.6", "ff.0 (h3)", "ff.2", "ff.4", "rgb_fn.0 + density_fn.0", "rgb_fn.2"]
off = 0
print(f"{'job':28s} ctas  total clk (mean / max)      MMA-done clk (mean)")
for n, k in zip(names, ctas):
    x = c[off:off + k]
    print(f"{n:28s} {k:3d}   {x[:,0].mean():10.0f} / {x[:,0].max():10.0f}   {x[:,1].mean():10.0f}")
    off += k
