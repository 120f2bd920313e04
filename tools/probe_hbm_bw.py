"""Write-only and read-only HBM bandwidth with plain 16-byte grid-stride kernels (the copy figure in MEASURED_PEAKS.json is
read + write).  The write-only number is the roofline of the training-form forward and the dgrad kernel."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import _native as nat
fn = nat.debug_lib().nerf_debug_hbm_bw
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
buf = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
for mode, name in ((0, "write-only"), (1, "read-only"), (2, "write, training-tensor pattern")):
    for bps in ((1, 2) if mode == 2 else (2, 4, 8)):
        for _ in range(2):
            nat.check(fn(nat.ptr(buf), buf.numel(), mode, bps, nat.ptr(sink), None), "probe")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            nat.check(fn(nat.ptr(buf), buf.numel(), mode, bps, nat.ptr(sink), None), "probe")
        b.record(); torch.cuda.synchronize()
        print(f"{name:32s} 8 GiB, {bps} x 512-thread blocks per SM: {3 * buf.numel() / a.elapsed_time(b) / 1e6:8.1f} GB/s")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
half = buf.numel() // 2
a.record()
for _ in range(3): buf[:half].copy_(buf[half:])
b.record(); torch.cuda.synchronize()
print(f"copy (read + write bytes): {3 * 2 * half / a.elapsed_time(b) / 1e6:8.1f} GB/s")
