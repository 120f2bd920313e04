"""Per-CTA cycle counters of the fused tcgen05 MLP kernel (diagnostic build of the same kernel).
usage: python tools/profile_mlp_tc.py [N] [S]"""
import ctypes
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import _native as nat
import nerf_model
import synthetic

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
dev = torch.device("cuda")
net = nerf_model.NeRFNetwork()
net.load_state_dict(synthetic.make_state_dict(5, "dense"))
net = net.to(dev)
m = net.fine_network
packed = m.packed_weights()
o = torch.randn(N, 3, device=dev) * 0.3
d = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=1)
ts = (2 + 4 * torch.rand(N, S, device=dev)).contiguous()
sigma = torch.empty(N, S, device=dev)
rgb = torch.empty(N, S, 3, device=dev)
if os.environ.get("NERF_PROF_PLAIN") == "1":
    # production kernel (no counters), 20 back-to-back launches: the number to compare kernel variants with
    plain = nat.lib().nerf_mlp_forward_tc
    for rep in range(3):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            nat.check(plain(nat.ptr(packed), nat.ptr(o), nat.ptr(d), nat.ptr(ts), N, S, nat.ptr(sigma), nat.ptr(rgb), nat.stream()), "mlp")
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 20
        print(f"plain kernel N={N} S={S}: {ms:.4f} ms, {N*S*920832/ms/1e9:.1f} TFLOP/s")
    # the same network with compositing fused into the kernel (render form: no per-sample outputs)
    fused = nat.lib().nerf_mlp_composite_tc
    ts_sorted = torch.sort(ts, dim=1).values.contiguous()
    col, depth, acc, stats = torch.empty(N, 3, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev), torch.zeros(4, device=dev)
    w = torch.empty(N, S, device=dev)
    for want_w in (False, True):
        for rep in range(3):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                nat.check(fused(nat.ptr(packed), nat.ptr(o), nat.ptr(d), nat.ptr(ts_sorted), N, S, None, None, None, None,
                                nat.ptr(w) if want_w else None, nat.ptr(col), nat.ptr(depth), nat.ptr(acc), nat.ptr(stats), nat.stream()), "fused")
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 20
            print(f"fused kernel (weights {'on' if want_w else 'off'}) N={N} S={S}: {ms:.4f} ms, {N*S*920832/ms/1e9:.1f} TFLOP/s")
    sys.exit(0)
fn = nat.debug_lib().nerf_debug_mlp_tc_profile
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int] + [ctypes.c_void_p] * 4
dbg = torch.zeros(148 * 16 + 128, dtype=torch.int64, device=dev)
import os
G = 148
for it in range(3):
    dbg.zero_()
    dbg[148 * 16 + 120] = int(os.environ.get('NERF_PROF_FLAGS', '0'))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    nat.check(fn(nat.ptr(packed), nat.ptr(o), nat.ptr(d), nat.ptr(ts), N, S, nat.ptr(sigma), nat.ptr(rgb), nat.ptr(dbg), nat.stream()), "profile")
    t1.record()
    torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
detail = dbg[148 * 16:].double().cpu()
c = dbg[:148 * 16].view(148, 16).double().cpu()[:G]
tiles = c[:, 8].clamp(min=1)
one_tile = False          # (the one-tile schedule was removed in round 2; its counter names are kept for old records)
names3 = ["mma_total", "mma_wait_full(weights)", "mma_wait_dfree(acc read)", "mma_wait_alo/ahi(epilogue)", "mma_wait_pe",
          "producer_wait_empty", "epiX_total", "epiX_wait_dfull", "pairs", "mma_in_issue_blocks(layer_half)", "epiX_h1: dfull->alo (x7)",
          "epiX_h1: dfull->dfree (x7)", "epiX_h1: dfull->ahi (x7)", "epiX_h0: dfull->packed (x7)"]
names = ["mma_total", "mma_wait_full(weights)", "mma_wait_edone(epilogue)", "mma_wait_pe", "producer_wait_empty", "epi_total",
         "epi_wait_dfull", "epi_pe_time", "tiles", "epi_seg_ld+wait", "epi_seg_math1", "epi_seg_wait_st", "epi_seg_fence+arrive", "epi_seg_st1", "epi_seg_math2(+act store)", "epi_seg_st2"]
if not one_tile:
    names = names3
    tiles = tiles * 2          # counter 8 holds tile PAIRS for the two-tile kernel
print(f"N={N} S={S}: {ms:.3f} ms, {N*S*920832/ms/1e9:.1f} TFLOP/s, tiles/CTA {tiles.mean():.1f}")
for i, n in enumerate(names):
    if i == 8: continue
    if i >= c.shape[1]: break
    per_tile = (c[:, i] / tiles)
    print(f"  {n:28s} per tile: mean {per_tile.mean():9.0f} clk  min {per_tile.min():9.0f}  max {per_tile.max():9.0f}")
print("  (tensor-pipe floor per tile: 14.7k clk)")

if not one_tile:
    pairs0 = float(dbg[8].item())
    d = detail / max(pairs0, 1)
    print("CTA 0, clk per pair: weight-stage waits by stage index:")
    print("  " + " ".join(f"{v:4.0f}" for v in d[:58].tolist()))
    print("dfree waits by step, tile X / tile Y:")
    print("  X " + " ".join(f"{v:5.0f}" for v in d[32:48].tolist()))
    print("  Y " + " ".join(f"{v:5.0f}" for v in d[48:64].tolist()))
    print("alo/ahi waits by step, tile X / tile Y:")
    print("  X " + " ".join(f"{v:5.0f}" for v in d[64:80].tolist()))
    print("  Y " + " ".join(f"{v:5.0f}" for v in d[80:96].tolist()))
    print(f"layer_half total per pair (13 calls): {d[95].item():.0f} clk")
