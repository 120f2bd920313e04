"""TMEM load/store bandwidth per SM (one CTA), for 4/8/16 warps."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import _native as nat
fn = nat.debug_lib().nerf_debug_tmem_bw
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(8, dtype=torch.int64, device="cuda")
for mode, name in ((0, "ld x32 (4 KB/warp-instr)"), (1, "st x16 (2 KB/warp-instr)"), (2, "ld x32 + st x16")):
    for nw in (1, 4, 8, 16):
        for _ in range(2):
            nat.check(fn(nw, 2000, mode, nat.ptr(out), None), "probe")
            torch.cuda.synchronize()
        cyc, ldbytes = out[0].item(), out[1].item()
        b = ldbytes if mode != 1 else ldbytes // 2
        print(f"{name:28s} warps={nw:2d}: {cyc/2000:8.1f} clk/iter/warp-set, {b/cyc:7.1f} B/clk/SM ({'load' if mode!=1 else 'store'} bytes)")
