"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/nerf_b200.h declares; the host modules keep the reference's call surface; no CPU fallback exists."""
import ctypes
import inspect
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import _native
    if not _native.LIB_PATH.exists():
        _native.build()
    return _native.lib()


def test_header_symbols_exported(lib):
    header = (ROOT / "include" / "nerf_b200.h").read_text()
    declared = set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 16
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/nerf_b200.h but not exported"
    import _native
    assert declared == set(_native.exported_symbols())
    assert lib.nerf_abi_version() == 3
    assert lib.nerf_packed_bytes() == 57 * 16384 + 6 * 2048 + 1928 * 4


def test_ctypes_prototypes_match_the_header():
    """Every declaration of include/nerf_b200.h has a ctypes prototype in _native._PROTOTYPES with the same number of arguments
    and a matching kind per argument (pointer / 64-bit integer / int / float): a signature edited on one side only would
    otherwise corrupt the call silently."""
    import _native
    header = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "nerf_b200.h").read_text(), flags=re.S)
    decls = re.findall(r"NERF_API\s+[\w\s\*]+?\b(nerf_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    assert len(decls) == len(_native._PROTOTYPES)

    def kind_of_c(arg):
        arg = " ".join(arg.split())
        if "*" in arg:
            return "pointer"
        return {"int64_t": "i64", "int": "int", "float": "float", "size_t": "size"}[arg.rsplit(" ", 1)[0].replace("const ", "")]

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) == "P":
            return "pointer"
        return {ctypes.c_int64: "i64", ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_size_t: "size"}[t]

    for name, args in decls:
        args = args.strip()
        c_kinds = [] if args in ("", "void") else [kind_of_c(a) for a in args.split(",")]
        proto = _native._PROTOTYPES[name][1]
        assert [kind_of_ctypes(t) for t in proto] == c_kinds, name


def test_dynamic_symbol_table_is_exactly_the_header(lib):
    """The product library is built with -fvisibility=hidden: `nm -D` must list the functions of include/nerf_b200.h and
    nothing else of ours (no nerf_debug_* probes, no internal C++ symbols); the probes live in tools/libnerf_b200_debug.so."""
    import subprocess
    import _native
    header = (ROOT / "include" / "nerf_b200.h").read_text()
    declared = set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", header))
    out = subprocess.run(["nm", "-D", "--defined-only", str(_native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    exported = {s for s in exported if not s.startswith(("_init", "_fini", "__bss_start", "_edata", "_end"))}
    assert exported == declared, (sorted(exported - declared), sorted(declared - exported))
    assert not any("debug" in s for s in exported)
    if _native.DEBUG_LIB_PATH.exists():
        dbg = subprocess.run(["nm", "-D", "--defined-only", str(_native.DEBUG_LIB_PATH)], capture_output=True, text=True, check=True).stdout
        assert "nerf_debug_umma" in dbg and "nerf_debug_mlp_tc_profile" in dbg


def test_argument_validation_without_gpu(lib):
    # null pointers are rejected before any CUDA call, with a message
    assert lib.nerf_deltas(None, 4, 4, None, None) == -1
    assert b"null pointer" in lib.nerf_last_error()
    assert lib.nerf_merge_sort(None, None, None, 1, None, 1, 1, None, None, None) == -1


def test_no_cpu_fallback():
    import nerf_helpers
    import nerf_model
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        nerf_helpers.generate_deltas(torch.zeros(1, 4, 1))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        nerf_model.NeRFModel()(torch.zeros(4, 4, 3), torch.zeros(4, 3))
    src = "".join(inspect.getsource(m) for m in (nerf_helpers, nerf_model))
    assert "oracle" not in src


def test_render_entry_point_surface():
    """render.py keeps the reference's call (render.py:14) and command line (render.py:20-25); the epoch tag is cut out of the
    checkpoint name the way upstream does it."""
    import render
    assert list(inspect.signature(render.render).parameters) == ["ckpt", "save_dir", "rays", "num_poses"]
    ns = render.build_parser().parse_args(["-c", "models/model=lego-epoch=1089-step=108999.ckpt"])
    assert (ns.rays, ns.num_poses, str(ns.save_dir)) == (4096, 40, "recons")
    assert render.epoch_tag(ns.ckpt) == "epoch=1089" and render.epoch_tag("x/epoch=3-step=11.ckpt") == "epoch=3"
    with pytest.raises(SystemExit):
        render.build_parser().parse_args([])                       # -c is required


def test_call_surface_matches_reference():
    """Names, positional parameters and defaults of the reference's public functions."""
    import dataloader
    import nerf_helpers as h
    import nerf_model as m

    def params(fn):
        return [(k, v.default) for k, v in inspect.signature(fn).parameters.items()]
    assert params(h.generate_coarse_samples)[:5] == [("o_rays", inspect._empty), ("d_rays", inspect._empty),
                                                      ("num_samples", inspect._empty), ("near", 2.0), ("far", 6.0)]
    assert [k for k, _ in params(h.inverse_transform_sampling)][:7] == ["o_rays", "d_rays", "weights", "ts", "num_samples", "near", "far"]
    assert [k for k, _ in params(h.view_reconstruction)] == ["model", "all_o_rays", "all_d_rays", "N"]
    assert params(h.generate_360_view_synthesis)[3:] == [("height", 800), ("width", 800), ("radius", 4.0),
                                                         ("cam_angle_x", 0.6911112070083618), ("N", 4096), ("num_poses", 40)]
    assert [k for k, _ in params(dataloader.get_rays)][:4] == ["H", "W", "focal", "c2w"]
    assert params(m.NeRFNetwork.__init__)[1:7] == [("position_dim", 10), ("direction_dim", 4), ("coarse_samples", 64),
                                                   ("fine_samples", 128), ("near", 2.0), ("far", 6.0)]
    import synthetic
    net = m.NeRFNetwork()
    assert list(net.state_dict().keys()) == synthetic.state_dict_keys()
    assert sum(p.numel() for p in net.parameters()) == 924680
