"""ctypes binding of libsplatter360.so (the C-ABI declared in include/splatter360.h).

The library is built in-tree by ``splatter360_b200.csrc.build`` (nvcc, sm_100a).  There is NO
fallback: if the shared object is missing or a symbol is absent, loading raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S360_LIB", os.path.join(_HERE, "libsplatter360.so"))  # S360_LIB: A/B-test a variant build

MODE_PINHOLE = 0
MODE_ERP = 1
DEPTH_MODES = {"depth": 0, "disparity": 1, "relative_disparity": 2, "log": 3}

EXPORTS = [
    "s360_abi_version", "s360_error_string", "s360_launch_count",
    "s360_geom_bytes", "s360_preprocess_scratch_bytes", "s360_binning_scratch_bytes",
    "s360_image_bytes", "s360_backward_scratch_bytes",
    "s360_forward_preprocess", "s360_forward_project", "s360_forward_order", "s360_forward_render", "s360_backward", "s360_mark_visible",
    "s360_debug_unpack_geom", "s360_debug_unpack_image",
    "s360_profile_enable", "s360_profile_read", "s360_mse_loss_grad",
    "s360_multi_geom_bytes", "s360_multi_preprocess_scratch_bytes", "s360_multi_binning_scratch_bytes",
    "s360_multi_image_bytes", "s360_multi_backward_scratch_bytes",
    "s360_multi_forward_project", "s360_multi_forward_order", "s360_multi_forward_render", "s360_multi_backward",
    "s360_debug_unpack_pairs", "s360_cube2equirec_forward", "s360_cube2equirec_backward", "s360_debug_counters",
    "s360_adapter_forward", "s360_adapter_backward", "s360_invert4x4",
]
ABI_VERSION = 6
MAX_VIEWS = 32

STAGES = ["preprocess", "depth_sort", "scan", "emit", "tile_sort", "tile_ranges", "render_fwd", "render_bwd",
          "preprocess_bwd"]


class S360View(ctypes.Structure):
    """Mirror of ``struct S360View`` (include/splatter360.h)."""
    _fields_ = [
        ("P", c_int32), ("M", c_int32), ("sh_degree", c_int32),
        ("image_height", c_int32), ("image_width", c_int32), ("mode", c_int32),
        ("max_sh_degree", c_int32), ("tight_bbox", c_int32),
        ("tanfovx", c_float), ("tanfovy", c_float), ("near_cull", c_float),
        ("fov_clamp", c_float), ("lowpass", c_float), ("pole_eps", c_float),
        ("scene_scale", c_float), ("sh_layout", c_int32), ("cov_layout", c_int32), ("reserved0", c_int32),
        ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("campos", c_void_p), ("bg", c_void_p),
    ]


_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (building it first if the sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from .csrc.build import build
            build()
        except Exception as e:  # pragma: no cover
            raise ImportError(
                f"libsplatter360.so not found at {LIB_PATH} and could not be built ({e}). "
                "Run `python -m splatter360_b200.csrc.build`. There is no CPU fallback.") from e
    lib = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise ImportError(f"libsplatter360.so does not export {name}")
    lib.s360_abi_version.restype = c_int
    lib.s360_error_string.restype = c_char_p
    lib.s360_error_string.argtypes = [c_int]
    lib.s360_launch_count.restype = c_uint64
    for n in ("s360_geom_bytes", "s360_preprocess_scratch_bytes", "s360_backward_scratch_bytes"):
        getattr(lib, n).restype = c_size_t
        getattr(lib, n).argtypes = [c_int32]
    lib.s360_binning_scratch_bytes.restype = c_size_t
    lib.s360_binning_scratch_bytes.argtypes = [c_int32, c_int64, c_int32, c_int32]
    lib.s360_image_bytes.restype = c_size_t
    lib.s360_image_bytes.argtypes = [c_int32, c_int32]
    vp = c_void_p
    lib.s360_forward_preprocess.restype = c_int
    lib.s360_forward_preprocess.argtypes = [ctypes.POINTER(S360View)] + [vp] * 12
    lib.s360_forward_project.restype = c_int
    lib.s360_forward_project.argtypes = [ctypes.POINTER(S360View)] + [vp] * 10
    lib.s360_forward_order.restype = c_int
    lib.s360_forward_order.argtypes = [ctypes.POINTER(S360View)] + [vp] * 6
    lib.s360_forward_render.restype = c_int
    lib.s360_forward_render.argtypes = [ctypes.POINTER(S360View), vp, vp, vp, vp, c_int64, vp, vp, vp, vp, c_int32, c_float, c_float, vp, vp]
    lib.s360_backward.restype = c_int
    lib.s360_backward.argtypes = [ctypes.POINTER(S360View)] + [vp] * 10 + [vp, c_int32, c_float, c_float] + [vp] * 8
    lib.s360_mark_visible.restype = c_int
    lib.s360_mark_visible.argtypes = [ctypes.POINTER(S360View), vp, vp, vp]
    lib.s360_debug_unpack_geom.restype = c_int
    lib.s360_debug_unpack_geom.argtypes = [c_int32] + [vp] * 8
    lib.s360_debug_unpack_image.restype = c_int
    lib.s360_debug_unpack_image.argtypes = [c_int32, c_int32] + [vp] * 5
    lib.s360_invert4x4.restype = c_int
    lib.s360_invert4x4.argtypes = [vp, vp, c_int64, vp]
    lib.s360_mse_loss_grad.restype = c_int
    lib.s360_mse_loss_grad.argtypes = [vp, vp, c_int64, c_float, vp, vp, vp]
    lib.s360_profile_enable.restype = c_int
    lib.s360_profile_enable.argtypes = [c_int]
    lib.s360_profile_read.restype = c_int
    lib.s360_profile_read.argtypes = [vp, vp, c_int]
    # batched multi-view path
    lib.s360_multi_geom_bytes.restype = c_size_t
    lib.s360_multi_geom_bytes.argtypes = [c_int32, c_int64]
    lib.s360_multi_preprocess_scratch_bytes.restype = c_size_t
    lib.s360_multi_preprocess_scratch_bytes.argtypes = [c_int32, c_int64]
    lib.s360_multi_binning_scratch_bytes.restype = c_size_t
    lib.s360_multi_binning_scratch_bytes.argtypes = [c_int64, c_int64, c_int32, c_int32, c_int32]
    lib.s360_multi_image_bytes.restype = c_size_t
    lib.s360_multi_image_bytes.argtypes = [c_int32, c_int32, c_int32]
    lib.s360_multi_backward_scratch_bytes.restype = c_size_t
    lib.s360_multi_backward_scratch_bytes.argtypes = [c_int64]
    head = [ctypes.POINTER(S360View), c_int32, c_int64]
    lib.s360_multi_forward_project.restype = c_int
    lib.s360_multi_forward_project.argtypes = head + [vp] * 10
    lib.s360_multi_forward_order.restype = c_int
    lib.s360_multi_forward_order.argtypes = head + [vp] * 6
    lib.s360_multi_forward_render.restype = c_int
    lib.s360_multi_forward_render.argtypes = head + [vp, vp, vp, vp, c_int64, vp, vp, vp, vp, c_int32, c_float, c_float, vp, vp]
    lib.s360_multi_backward.restype = c_int
    lib.s360_multi_backward.argtypes = head + [vp] * 9 + [vp, c_int32, c_float, c_float] + [vp] * 7
    for n in ("s360_cube2equirec_forward", "s360_cube2equirec_backward"):
        getattr(lib, n).restype = c_int
        getattr(lib, n).argtypes = [vp, vp] + [c_int32] * 6 + [vp, vp, vp]
    lib.s360_debug_unpack_pairs.restype = c_int
    lib.s360_debug_unpack_pairs.argtypes = [c_int32, c_int64] + [vp] * 5
    lib.s360_adapter_forward.restype = c_int
    lib.s360_adapter_forward.argtypes = [c_int32, c_int32, c_int32, c_int32, c_float, c_float] + [vp] * 10
    lib.s360_adapter_backward.restype = c_int
    lib.s360_adapter_backward.argtypes = [c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_int32] + [vp] * 10
    lib.s360_debug_counters.restype = c_int
    lib.s360_debug_counters.argtypes = [vp, c_int, vp]
    if lib.s360_abi_version() != ABI_VERSION:
        raise ImportError("libsplatter360.so ABI version mismatch")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        msg = load().s360_error_string(code).decode()
        raise RuntimeError(f"libsplatter360 error {code}: {msg}")


def profile_enable(on: bool) -> bool:
    return bool(load().s360_profile_enable(int(on)))


def profile_read(reset: bool = True) -> dict:
    """{stage: (total_ms, count)} since the last reset (synchronises on the recorded events)."""
    n = len(STAGES)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_uint64 * n)()
    check(load().s360_profile_read(ctypes.cast(ms, c_void_p), ctypes.cast(cnt, c_void_p), int(reset)))
    return {STAGES[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}


def launch_count() -> int:
    return int(load().s360_launch_count())
