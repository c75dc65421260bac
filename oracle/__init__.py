"""CPU oracle package -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``oracle.render`` wraps ``raster_oracle.c`` (a plain-C restatement of the rasterizer that
splatter360 reaches through ``diff_gaussian_rasterization``,
/root/reference/src/model/decoder/cuda_splatting.py:99-126).  ``oracle.torch_oracle`` is an
independent float64 autograd restatement used to pin the C oracle's gradients.

PARITY UNPINNED: the upstream CUDA extension is an un-vendored, un-pinned pip dependency
(requirements.txt:17) and the reference holds no golden vectors for this path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  ``splatter360_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "raster_oracle.c")
_SO = os.path.join(_HERE, "liboracle.so")


class OrcCfg(ctypes.Structure):
    _fields_ = [
        ("P", ctypes.c_int32),
        ("M", ctypes.c_int32),
        ("D", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("W", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("max_sh_degree", ctypes.c_int32),
        ("use_sh", ctypes.c_int32),
        ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float),
        ("near_cull", ctypes.c_float),
        ("fov_clamp", ctypes.c_float),
        ("lowpass", ctypes.c_float),
        ("pole_eps", ctypes.c_float),
        ("view", ctypes.c_float * 16),
        ("proj", ctypes.c_float * 16),
        ("campos", ctypes.c_float * 3),
        ("bg", ctypes.c_float * 3),
    ]


class OrcCfg64(ctypes.Structure):
    """OrcCfg of the float64 build (-DORACLE_F64): every float field is a double."""
    _fields_ = [(n, {ctypes.c_float: ctypes.c_double, ctypes.c_float * 16: ctypes.c_double * 16,
                     ctypes.c_float * 3: ctypes.c_double * 3}.get(t, t)) for n, t in OrcCfg._fields_]


_SO64 = os.path.join(_HERE, "liboracle_f64.so")
_lib64 = None


def lib64():
    """The float64 build of the same source: the reference both float32 paths are measured against."""
    global _lib64
    if _lib64 is None:
        if not os.path.exists(_SO64) or os.path.getmtime(_SO64) < os.path.getmtime(_SRC):
            subprocess.run(["gcc", "-O2", "-DORACLE_F64", "-fopenmp", "-shared", "-fPIC", "-o", _SO64, _SRC, "-lm"], check=True)
        _lib64 = ctypes.CDLL(_SO64)
        _lib64.oracle_render_ex.restype = ctypes.c_int
        _lib64.oracle_cfg_size.restype = ctypes.c_int
        assert _lib64.oracle_cfg_size() == ctypes.sizeof(OrcCfg64), "OrcCfg64 layout mismatch"
    return _lib64


def build(force: bool = False) -> str:
    """Compile raster_oracle.c -> liboracle.so and its float64 build liboracle_f64.so (gcc, OpenMP).  Returns the .so path."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    if force or not os.path.exists(_SO64) or os.path.getmtime(_SO64) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-DORACLE_F64", "-fopenmp", "-shared", "-fPIC", "-o", _SO64, _SRC, "-lm"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_render.restype = ctypes.c_int
        _lib.oracle_render_ex.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_cfg_size.restype = ctypes.c_int
        assert _lib.oracle_cfg_size() == ctypes.sizeof(OrcCfg), "OrcCfg layout mismatch"
    return _lib


DEFAULTS = dict(near_cull=0.2, fov_clamp=1.3, lowpass=0.3, pole_eps=1e-3, max_sh_degree=4)


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(n))


def render(
    means,
    cov6,
    opac,
    *,
    shs=None,
    colors=None,
    H: int,
    W: int,
    view,
    proj,
    campos,
    bg=(0.0, 0.0, 0.0),
    tanfovx: float = 1.0,
    tanfovy: float = 1.0,
    sh_degree: int = 0,
    mode: str = "pinhole",
    dL_dpix=None,
    stages: bool = True,
    depth_mode: Optional[str] = None,
    depth_near: float = 0.0,
    depth_far: float = 0.0,
    depth_scale: float = 1.0,
    dL_ddepth=None,
    f64: bool = False,
    **consts,
) -> dict:
    """Run the C oracle on one view.  All arrays numpy float32.

    means [P,3]; cov6 [P,6] (xx,xy,xz,yy,yz,zz); opac [P]; shs [P,M,3] or colors [P,3];
    view/proj: 4x4 as handed to GaussianRasterizationSettings (row-vector convention,
    cuda_splatting.py:85-87).  Returns dict with 'color' [3,H,W], 'radii', and (if
    stages) the per-stage intermediates; gradients if dL_dpix [3,H,W] is given.

    ``depth_mode`` ("depth" | "disparity" | "relative_disparity" | "log"): also blend the per-Gaussian depth value
    (of sort depth / depth_scale) as a fourth channel -> 'depth_image' [H,W]; ``dL_ddepth`` [H,W] adds its gradient to the
    returned gradients (the reference's second rasterisation with depth as colour, cuda_splatting.py:226-269).
    """
    # f64: the float64 build of the same source (lib64) -- inputs are the same float32 values, widened
    L = lib64() if f64 else lib()
    FT = np.float64 if f64 else np.float32
    cfloat = ctypes.c_double if f64 else ctypes.c_float

    def _f32(a, shape=None):   # noqa: F811  (conversion to the build's real type)
        a = np.ascontiguousarray(np.asarray(np.asarray(a, dtype=FT), dtype=FT))
        return a if shape is None else a.reshape(shape)

    k = dict(DEFAULTS)
    k.update(consts)
    means = _f32(means).reshape(-1, 3)
    P = means.shape[0]
    cov6 = _f32(cov6).reshape(P, 6)
    opac = _f32(opac).reshape(P)
    use_sh = shs is not None
    if use_sh:
        shs = _f32(shs)
        M = shs.shape[1]
        assert shs.shape == (P, M, 3)
    else:
        M = 0
        colors = _f32(colors).reshape(P, 3)
    cfg = OrcCfg64() if f64 else OrcCfg()
    cfg.P, cfg.M, cfg.D, cfg.H, cfg.W = P, M, int(sh_degree), int(H), int(W)
    cfg.mode = {"pinhole": 0, "erp": 1}[mode]
    cfg.max_sh_degree = int(k["max_sh_degree"])
    cfg.use_sh = int(use_sh)
    cfg.tanfovx, cfg.tanfovy = float(tanfovx), float(tanfovy)
    cfg.near_cull, cfg.fov_clamp = float(k["near_cull"]), float(k["fov_clamp"])
    cfg.lowpass, cfg.pole_eps = float(k["lowpass"]), float(k["pole_eps"])
    cfg.view[:] = _f32(view).reshape(16).tolist()
    cfg.proj[:] = _f32(proj).reshape(16).tolist()
    cfg.campos[:] = _f32(campos).reshape(3).tolist()
    cfg.bg[:] = _f32(bg).reshape(3).tolist()

    gx, gy = (W + 15) // 16, (H + 15) // 16
    out = {
        "color": np.zeros((3, H, W), FT),
        "radii": np.zeros(P, np.int32),
    }
    st = {}
    if stages or dL_dpix is not None:
        st = {
            "final_T": np.zeros((H, W), FT),
            "n_contrib": np.zeros((H, W), np.uint32),
            "xy": np.zeros((P, 2), FT),
            "depth": np.zeros(P, FT),
            "conic_opacity": np.zeros((P, 4), FT),
            "rgb": np.zeros((P, 3), FT),
            "tiles_touched": np.zeros(P, np.uint32),
            "clamped": np.zeros((P, 3), np.uint8),
            "tile_ranges": np.zeros((gx * gy, 2), np.uint32),
        }
    n_out = ctypes.c_int64(0)
    grads = {}
    dmode = -1 if depth_mode is None else {"depth": 0, "disparity": 1, "relative_disparity": 2, "log": 3}[depth_mode]
    if dmode >= 0:
        out["depth_image"] = np.zeros((H, W), FT)
    if dL_ddepth is not None:
        dL_ddepth = _f32(dL_ddepth).reshape(H, W)
        if dL_dpix is None:
            dL_dpix = np.zeros((3, H, W), FT)
    if dL_dpix is not None:
        dL_dpix = _f32(dL_dpix).reshape(3, H, W)
        grads = {
            "d_means": np.zeros((P, 3), FT),
            "d_means2D": np.zeros((P, 3), FT),
            "d_cov6": np.zeros((P, 6), FT),
            "d_opac": np.zeros(P, FT),
            "d_shs": np.zeros((P, M, 3), FT) if use_sh else None,
            "d_colors": np.zeros((P, 3), FT),
        }

    def call(cap, it, ig):
        return L.oracle_render_ex(
            ctypes.byref(cfg), _ptr(means), _ptr(cov6), _ptr(opac), _ptr(shs), _ptr(colors),
            _ptr(out["color"]), _ptr(out["radii"]), _ptr(st.get("final_T")), _ptr(st.get("n_contrib")),
            _ptr(st.get("xy")), _ptr(st.get("depth")), _ptr(st.get("conic_opacity")), _ptr(st.get("rgb")),
            _ptr(st.get("tiles_touched")), _ptr(st.get("clamped")),
            ctypes.c_int64(cap), _ptr(it), _ptr(ig), _ptr(st.get("tile_ranges")),
            ctypes.byref(n_out),
            _ptr(dL_dpix), _ptr(grads.get("d_means")), _ptr(grads.get("d_means2D")),
            _ptr(grads.get("d_cov6")), _ptr(grads.get("d_opac")), _ptr(grads.get("d_shs")),
            _ptr(grads.get("d_colors")),
            ctypes.c_int(dmode), cfloat(1.0 / depth_scale), cfloat(depth_near), cfloat(depth_far),
            _ptr(out.get("depth_image")), _ptr(dL_ddepth),
        )

    rc = call(0, None, None)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed rc={rc}")
    N = int(n_out.value)
    out["num_rendered"] = N
    if stages:
        it = np.zeros(max(N, 1), np.uint32)
        ig = np.zeros(max(N, 1), np.uint32)
        rc = call(N, it, ig)
        if rc != 0:
            raise RuntimeError(f"oracle_render failed rc={rc}")
        st["inst_tile"], st["inst_gid"] = it[:N], ig[:N]
    out.update(st)
    out.update({k_: v for k_, v in grads.items() if v is not None})
    return out
