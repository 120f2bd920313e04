"""GPU known-answer probe of the tcgen05 building blocks (umma.cuh): one-CTA GEMMs with A in swizzled shared
memory (SS) or in TMEM (TS) against torch.matmul on the same bf16 inputs.  The probe kernels live in the diagnostic
library (tools/libnerf_b200_debug.so, csrc/debug/umma_probe.cu), not in the product library."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K,N,d_col", [(64, 128, 0), (256, 128, 128), (256, 16, 128), (128, 256, 0), (64, 144, 0)])
def test_umma_probe(mode, K, N, d_col):
    import _native as nat
    fn = nat.debug_lib().nerf_debug_umma
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                   ctypes.c_void_p]
    g = torch.Generator(device="cuda").manual_seed(1234 + K + N)
    A = (torch.randn(128, K, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    B = (torch.randn(N, K, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    nat.check(fn(mode, nat.ptr(A), nat.ptr(B), K, N, d_col, nat.ptr(D), nat.stream()), "nerf_debug_umma")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = (D - ref).abs().max().item()
    assert err < 1e-2 * max(1.0, ref.abs().max().item() / 16), f"mode {mode} K {K} N {N}: max err {err}"


@pytest.mark.parametrize("K,N,d_col", [(128, 128, 0), (128, 256, 0), (64, 64, 64)])
def test_umma_probe_mn_major(K, N, d_col):
    """Both operands MN-major without swizzle (the wgrad form: contraction over samples): D = A^T B with A [K,128], B [K,N]."""
    import _native as nat
    fn = nat.debug_lib().nerf_debug_umma
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                   ctypes.c_void_p]
    g = torch.Generator(device="cuda").manual_seed(99 + K + N)
    A = torch.randn(K, 128, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    B = torch.randn(K, N, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    nat.check(fn(2, nat.ptr(A), nat.ptr(B), K, N, d_col, nat.ptr(D), nat.stream()), "nerf_debug_umma")
    torch.cuda.synchronize()
    ref = A.float().T @ B.float()
    err = (D - ref).abs().max().item()
    assert err < 1e-2 * max(1.0, ref.abs().max().item() / 16), f"MN-major K {K} N {N}: max err {err}"
