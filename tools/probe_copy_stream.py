"""Per-SM global(L2)->shared streaming throughput of cp.async.bulk vs tensor-map TMA through an mbarrier ring."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import _native as nat
fn = nat.debug_lib().nerf_debug_copy_stream
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
               ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
src = torch.randint(0, 255, (57 * 16384,), dtype=torch.uint8, device="cuda")
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for grid in (148,):
    for mode, name in ((0, "cp.async.bulk 1-D"), (1, "tensor-map TMA 2-D")):
        for nprod in (1, 2, 4):
            for nbytes, slots in ((16384, 9), (32768, 6), (65536, 3), (8192, 16)):
                if mode == 1 and nbytes > 32768:
                    continue
                nst = 4000
                for _ in range(2):
                    nat.check(fn(nat.ptr(src), src.numel(), mode, nbytes, slots, nst, grid, nprod, nat.ptr(out), None), "probe")
                    torch.cuda.synchronize()
                cyc = out[:grid].double().mean().item()
                print(f"{name:20s} issuers {nprod}  stage {nbytes:6d} B x {slots:2d} slots: {nbytes*nst/cyc:6.1f} B/clk/SM  ({cyc/nst:6.0f} clk/stage)")
