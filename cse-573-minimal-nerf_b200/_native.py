"""ctypes binding of libnerf_b200.so (include/nerf_b200.h).

There is no CPU implementation behind this module: loading fails loudly when the library has not been
built, and every wrapper rejects non-CUDA tensors.
"""
import ctypes
import subprocess
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libnerf_b200.so"
DEBUG_LIB_PATH = HERE.parent / "tools" / "libnerf_b200_debug.so"      # probes + cycle-counter kernel forms; tools/ and tests only
ABI_VERSION = 3

_c_f32p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_vp = ctypes.c_void_p

_PROTOTYPES = {
    "nerf_abi_version": (ctypes.c_int, []),
    "nerf_last_error": (ctypes.c_char_p, []),
    "nerf_raygen": (_int, [_vp, _int, _int, _f32, _vp, _vp, _i64, _vp, _vp, _vp]),
    "nerf_batch_rays": (_int, [_vp, _vp, _vp, _int, _int, _int, _f32, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "nerf_coarse_sample": (_int, [_vp, _vp, _vp, _vp, _f32, _i64, _int, _vp, _vp, _vp]),
    "nerf_deltas": (_int, [_vp, _i64, _int, _vp, _vp]),
    "nerf_weights": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "nerf_ray_color": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "nerf_composite": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerf_composite_backward": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp]),
    "nerf_fine_sample": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _f32, _f32, _vp, _vp, _vp, _vp]),
    "nerf_merge_sort": (_int, [_vp, _vp, _vp, _int, _vp, _int, _i64, _vp, _vp, _vp]),
    "nerf_fine_sample_merge": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _f32, _f32, _vp, _vp]),
    "nerf_positional_encoding": (_int, [_vp, _i64, _int, _int, _vp, _vp]),
    "nerf_mlp_forward_fp32": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _vp, _vp, _vp]),
    "nerf_packed_bytes": (ctypes.c_size_t, []),
    "nerf_pack_weights": (_int, [_vp, _vp, _vp]),
    "nerf_pack_weights_all": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "nerf_mlp_forward_tc": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp]),
    "nerf_mlp_forward_tc_train": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    "nerf_packed_t_bytes": (ctypes.c_size_t, []),
    "nerf_pack_weights_t": (_int, [_vp, _vp, _vp]),
    "nerf_mlp_backward_tc": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "nerf_mlp_backward_tc_fused": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "nerf_wgrad_tc": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _int, _vp]),
    "nerf_mlp_forward_tc_points": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp]),
    "nerf_mlp_composite_tc_supported": (_int, [_int]),
    "nerf_mlp_composite_tc": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerf_mlp_composite_tc_strata": (_int, [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerf_render_workspace_bytes": (ctypes.c_size_t, [_i64, _int, _int]),
    "nerf_render_forward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _i64, _int, _int, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerf_adam_step_dev": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "nerf_adam_step": (_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i64, _f32, _vp]),
}

_lib = None


def build(verbose=False, target="all"):
    """Compile csrc/*.cu for sm_100a into libnerf_b200.so, and the diagnostic library (same sources with
    -DNERF_DEBUG_BUILD + csrc/debug/*.cu) into tools/libnerf_b200_debug.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(HERE / "csrc"), "-j8", target]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libnerf_b200.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded library.  Raises if it is missing - there is deliberately no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the NeRF hot path has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)            # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        if handle.nerf_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libnerf_b200.so ABI {handle.nerf_abi_version()} != expected {ABI_VERSION}")
        _lib = handle
    return _lib


def exported_symbols():
    return list(_PROTOTYPES)


_debug_lib = None


def debug_lib():
    """tools/libnerf_b200_debug.so: the product sources built with -DNERF_DEBUG_BUILD plus the tcgen05 / TMEM / copy-engine /
    HBM probes (`nerf_debug_*`).  Never loaded by the product path; callers set restype / argtypes themselves."""
    global _debug_lib
    if _debug_lib is None:
        if not DEBUG_LIB_PATH.exists():
            raise RuntimeError(f"{DEBUG_LIB_PATH} not found: build it with `make -C {HERE / 'csrc'} debug`")
        _debug_lib = ctypes.CDLL(str(DEBUG_LIB_PATH))
    return _debug_lib


launches = 0          # kernels launched through this module (bench.py reports it as gpu_launches)
kernel_events = None  # when a list: (name, start_event, end_event) appended around each timed tensor-core launch


def check(code, what):
    global launches
    if code != 0:
        raise RuntimeError(f"{what} failed ({code}): {lib().nerf_last_error().decode()}")
    launches += 1


class timed_kernel:
    """Brackets one kernel launch with CUDA events on the current stream when `kernel_events` is a list."""

    def __init__(self, name, units):
        self.name, self.units = name, units

    def __enter__(self):
        if kernel_events is not None:
            self.t0 = torch.cuda.Event(enable_timing=True)
            self.t1 = torch.cuda.Event(enable_timing=True)
            self.t0.record()
        return self

    def __exit__(self, *exc):
        if kernel_events is not None:
            self.t1.record()
            kernel_events.append((self.name, self.units, self.t0, self.t1))
        return False


def dev(t, name, dtype=torch.float32):
    """Validate a tensor argument: CUDA, expected dtype, contiguous.  Returns the (possibly re-laid-out) tensor."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this path has no CPU implementation), got device {t.device}")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def device_guard(fn):
    """Run `fn` with the CUDA device of its first CUDA-tensor argument (or of `self`'s parameters) current: the wrappers launch
    on `torch.cuda.current_stream()`, and a kernel for tensors on cuda:1 must not be queued while cuda:0 is the current device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dv = None
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda:
                    dv = a.device
                    break
            elif isinstance(a, torch.nn.Module):
                p = next(a.parameters(), None)
                if p is not None and p.is_cuda:
                    dv = p.device
        if dv is None or dv.index is None or dv.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dv):
            return fn(*args, **kwargs)
    return wrapper


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
