"""Does tcgen05.ld / st traffic slow down tcgen05.mma (and vice versa)?  One CTA: warp 0 streams MMAs, nw warps loop over
TMEM loads on a chosen column range.  Prints clk per MMA and the TMEM load rate for each placement."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import _native as nat
fn = nat.debug_lib().nerf_debug_tmem_contention
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 10 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(8, dtype=torch.int64, device="cuda")
NM = 4000

def run(N, d_col, a_col, nw, mode, ld_col, ld_span, commit_every=0, nm=NM, alt_every=0):
    for _ in range(2):
        out.zero_()
        nat.check(fn(nm, N, d_col, a_col, nw, mode, ld_col, ld_span, commit_every, alt_every, nat.ptr(out), None), "probe")
        torch.cuda.synchronize()
    cyc, ldb, stb, iss = (out[i].item() for i in range(4))
    return cyc / nm, ldb / cyc, stb / cyc, iss / nm

if len(sys.argv) > 1 and sys.argv[1] == "queue":
    print("queue depth: n_mma  issue clk (total)  completion clk (total)   [N=128, A in TMEM]")
    for nm in (1, 2, 4, 8, 16, 32, 64, 128):
        c, _, _, i = run(128, 0, 256, 0, 0, 0, 32, 0, nm)
        print(f"  {nm:4d} {i*nm:10.0f} {c*nm:10.0f}")
    print("commit cost: commit_every  clk/MMA  (4000 MMAs, N=128, A in TMEM)")
    for ce in (0, 32, 16, 8, 4, 2, 1):
        c, _, _, i = run(128, 0, 256, 0, 0, 0, 32, ce)
        print(f"  {ce:4d} {c:8.1f}")
    print("accumulator switching: alt_every  clk/MMA (N=128, A in TMEM; 16 epilogue-like ld+st warps running)")
    for ae in (0, 64, 16, 8, 4, 1):
        c, _, _, i = run(128, 0, 256, 0, 0, 0, 32, 0, NM, ae)
        c2, l2, s2, _ = run(128, 0, 256, 16, 2, 0, 256, 0, NM, ae)
        print(f"  {ae:4d} {c:8.1f}   with 16 ld+st warps: {c2:8.1f} (ld {l2:.0f} B/clk, st {s2:.0f} B/clk)")
    sys.exit(0)

print("N d_col a_src   ld_warps mode  ld_cols      clk/MMA  issue clk/MMA  ld B/clk  st B/clk")
for N in (128, 256):
    for a_col, a_name in ((-1, "smem"), (256, "tmem256")):
        for nw, mode, ld_col, span, tag in ((0, 0, 0, 32, "-"), (8, 1, 0, 128, "same D"), (8, 1, 128, 128, "D+128"), (8, 1, 256, 128, "A cols"),
                                            (16, 1, 0, 128, "same D"), (16, 1, 128, 128, "D+128"), (16, 1, 384, 64, "384.."),
                                            (8, 2, 128, 128, "D+128"), (16, 2, 128, 128, "D+128"), (4, 1, 128, 128, "D+128")):
            if N == 256 and ld_col == 128: ld_col = 256
            c, l, s_, i = run(N, 0, a_col, nw, mode, ld_col, span)
            print(f"{N:3d} {0:5d} {a_name:8s} {nw:5d}   {('none','ld','ld+st')[mode]:5s} {tag:10s} {c:9.1f} {i:12.1f} {l:9.1f} {s_:9.1f}")
