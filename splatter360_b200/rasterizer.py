"""Python surface of the rasterizer -- same names, arguments and error behaviour as the
``diff_gaussian_rasterization`` package the reference imports
(/root/reference/src/model/decoder/cuda_splatting.py:5-8) and calls
(/root/reference/src/model/decoder/cuda_splatting.py:99-126, 178-219):

    GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg, scale_modifier,
                                  viewmatrix, projmatrix, sh_degree, campos, prefiltered, debug)
    GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs=None, colors_precomp=None,
                                        scales=None, rotations=None, cov3D_precomp=None) -> (color, radii)

Everything numeric runs in libsplatter360.so (hand-written sm_100a CUDA, C-ABI in
include/splatter360.h) on the caller's current CUDA stream.  There is no CPU path: CPU tensors or a
missing library raise.

Extensions over upstream are optional trailing settings fields with defaults, so reference code that
constructs the settings by keyword keeps working unchanged:
``projection`` ("pinhole" | "erp"), ``near_cull``, ``fov_clamp``, ``lowpass``, ``pole_eps``,
``max_sh_degree``, ``tight_bbox``.
"""
from __future__ import annotations

import ctypes
import threading
from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool
    # ---- extensions (defaults reproduce upstream behaviour) ----
    projection: str = "pinhole"   # "pinhole" (upstream) or "erp" (native equirectangular)
    near_cull: float = 0.2
    fov_clamp: float = 1.3
    lowpass: float = 0.3
    pole_eps: float = 1e-3
    max_sh_degree: int = 4
    tight_bbox: bool = True
    scene_scale: float = 1.0      # means *= s, cov3D *= s^2 inside the kernels (cuda_splatting.py:64-71), grads w.r.t. inputs
    sh_layout: int = 0            # 0: shs [P,M,3]   1: [P,3,M] (the reference's harmonics layout, no transpose copy)
    cov_layout: int = 0           # 0: cov3D [P,6]   1: [P,3,3] (upper triangle read; gradient in the same layout)
    instance_capacity: Optional[int] = None   # None: read the instance count back (exact buffers, one small host read);
    #                                           int: trust this capacity -> no host read at all (CUDA-graph capturable); the
    #                                           device-side overflow flag is in ForwardState.counters, see overflowed()
    depth_mode: Optional[str] = None   # fused, differentiable depth channel: "depth" | "disparity" | "relative_disparity" | "log"
    depth_near: float = 0.0            # unscaled near / far used by relative_disparity and log
    depth_far: float = 0.0
    pair_capacity: Optional[int] = None   # batched path only: slots for (view, Gaussian) pairs; None = V * P (never overflows)
    capacity_tracker: Optional[object] = None   # a CapacityTracker: sync-free instance_capacity learnt from earlier calls


_MODES = {"pinhole": _lib.MODE_PINHOLE, "erp": _lib.MODE_ERP}


def _f32c(t: Tensor, device) -> Tensor:
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[Tensor]):
    return None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _make_view(s: GaussianRasterizationSettings, P: int, M: int, device, views: int = 1):
    """Build the host S360View plus the device tensors it points at (returned to keep them alive).  ``views`` > 1:
    viewmatrix / projmatrix / campos hold that many consecutive cameras (batched path)."""
    if s.projection not in _MODES:
        raise ValueError(f"unknown projection {s.projection!r} (expected 'pinhole' or 'erp')")
    if s.projection == "erp" and int(s.image_width) % 16 != 0:
        raise ValueError("erp projection needs image_width to be a multiple of 16 (seam wrap at tile granularity)")
    vm = _f32c(torch.as_tensor(s.viewmatrix), device)
    pm = _f32c(torch.as_tensor(s.projmatrix), device)
    cp = _f32c(torch.as_tensor(s.campos), device)
    bg = _f32c(torch.as_tensor(s.bg), device)
    if vm.numel() != 16 * views or pm.numel() != 16 * views or cp.numel() != 3 * views or bg.numel() != 3:
        raise ValueError("viewmatrix/projmatrix must have 16 elements (per view), campos 3 (per view), bg 3")
    v = _lib.S360View()
    v.P, v.M, v.sh_degree = int(P), int(M), int(s.sh_degree)
    v.image_height, v.image_width = int(s.image_height), int(s.image_width)
    v.mode = _MODES[s.projection]
    v.max_sh_degree = int(s.max_sh_degree)
    v.tight_bbox = int(bool(s.tight_bbox))
    v.tanfovx, v.tanfovy = float(s.tanfovx), float(s.tanfovy)
    v.near_cull, v.fov_clamp = float(s.near_cull), float(s.fov_clamp)
    v.lowpass, v.pole_eps = float(s.lowpass), float(s.pole_eps)
    v.scene_scale, v.sh_layout, v.cov_layout = float(s.scene_scale), int(s.sh_layout), int(s.cov_layout)
    v.viewmatrix, v.projmatrix = vm.data_ptr(), pm.data_ptr()
    v.campos, v.bg = cp.data_ptr(), bg.data_ptr()
    return v, (vm, pm, cp, bg)


def build_covariance_6(scales: Tensor, rotations: Tensor, scale_modifier: float = 1.0) -> Tensor:
    """cov3D[P,6] = R S S^T R^T from scales [P,3] and (r,x,y,z) quaternions [P,4] (upstream computeCov3D,
    SURVEY.md sec. 2b).  Differentiable torch ops; the kernels always consume the 6-vector."""
    s = scales * scale_modifier
    r, x, y, z = rotations.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    Mx = R * s[:, None, :]
    cov = Mx @ Mx.transpose(1, 2)
    return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], -1)


_readers = {}
_readers_lock = threading.Lock()


def _count_reader(device):
    """Per-device (pinned 16-byte buffer, side stream, event, lock) used to read S360Counters back early.  The lock
    serialises the read-back of concurrent forward calls on the same device (host threads share the pinned buffer)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    with _readers_lock:
        r = _readers.get(key)
        if r is None:
            r = (torch.empty(4, dtype=torch.int32).pin_memory(), torch.cuda.Stream(device=device), torch.cuda.Event(),
                 threading.Lock())
            _readers[key] = r
    return r


class ForwardState(NamedTuple):
    """Buffers kept from forward to backward (all caller-owned torch tensors)."""
    geom: Tensor
    radii: Tensor
    point_list: Tensor
    image_state: Tensor
    num_rendered: int
    num_visible: int
    depth: Optional[Tensor] = None   # [H,W] fused depth channel when settings.depth_mode is set
    counters: Optional[Tensor] = None   # device int32[4]: num_rendered, overflow, num_visible, -


def forward_raw(settings: GaussianRasterizationSettings, means3D: Tensor, cov6: Tensor, opacities: Tensor,
                shs: Optional[Tensor], colors: Optional[Tensor]):
    """Run the two forward stages through the C-ABI.  Returns (color[3,H,W], ForwardState)."""
    lib = _lib.load()
    device = means3D.device
    if device.type != "cuda":
        raise RuntimeError("splatter360_b200 rasterizer needs CUDA tensors (there is no CPU path)")
    P = means3D.shape[0]
    M = (shs.shape[2] if settings.sh_layout else shs.shape[1]) if shs is not None else 0
    H, W = int(settings.image_height), int(settings.image_width)
    tracker = settings.capacity_tracker
    if tracker is not None and settings.instance_capacity is None:
        settings = tracker.settings(settings)   # capacity from earlier calls once a count has arrived; else exact path
    with torch.cuda.device(device):
        view, keep = _make_view(settings, P, M, device)
        u8 = dict(dtype=torch.uint8, device=device)
        geom = torch.empty(lib.s360_geom_bytes(P), **u8)
        pre_scratch = torch.empty(lib.s360_preprocess_scratch_bytes(P), **u8)
        radii = torch.empty(P, dtype=torch.int32, device=device)
        depth_order = torch.empty(max(P, 1), dtype=torch.int32, device=device)
        offsets = torch.empty(max(P, 1), dtype=torch.int32, device=device)
        counters = torch.empty(4, dtype=torch.int32, device=device)
        st = _stream_ptr()
        # K1 fixes the instance count; it is read back over a side stream while the depth sort and scan run, so
        # the data-dependent allocation below does not leave the GPU idle
        _lib.check(lib.s360_forward_project(
            ctypes.byref(view), _ptr(means3D), _ptr(cov6), _ptr(opacities), _ptr(shs), _ptr(colors),
            _ptr(geom), _ptr(radii), _ptr(counters), _ptr(pre_scratch), st))
        if settings.instance_capacity is not None:
            # sync-free path: nothing is read back; num_rendered / num_visible stay on the device
            _lib.check(lib.s360_forward_order(
                ctypes.byref(view), _ptr(geom), _ptr(depth_order), _ptr(offsets), _ptr(counters), _ptr(pre_scratch), st))
            N, nvis = int(settings.instance_capacity), -1
        else:
            host_counts, side, ready, lock = _count_reader(device)
            with lock:
                ready.record(torch.cuda.current_stream())
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    host_counts.copy_(counters, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(side)
                counters.record_stream(side)
                _lib.check(lib.s360_forward_order(
                    ctypes.byref(view), _ptr(geom), _ptr(depth_order), _ptr(offsets), _ptr(counters), _ptr(pre_scratch), st))
                done.synchronize()
                N, nvis = int(host_counts[0].item()) & 0xFFFFFFFF, int(host_counts[2].item()) & 0xFFFFFFFF
        cap = max(N, 1)
        point_list = torch.empty(cap, dtype=torch.int32, device=device)
        bin_scratch = torch.empty(lib.s360_binning_scratch_bytes(P, cap, H, W), **u8)
        image_state = torch.empty(lib.s360_image_bytes(H, W), **u8)
        color = torch.empty((3, H, W), dtype=torch.float32, device=device)
        depth = None
        dmode = 0
        if settings.depth_mode is not None:
            if settings.depth_mode not in _lib.DEPTH_MODES:
                raise ValueError(f"unknown depth_mode {settings.depth_mode!r}")
            dmode = _lib.DEPTH_MODES[settings.depth_mode]
            depth = torch.empty((H, W), dtype=torch.float32, device=device)
        _lib.check(lib.s360_forward_render(
            ctypes.byref(view), _ptr(geom), _ptr(depth_order), _ptr(offsets), _ptr(counters),
            ctypes.c_int64(N), _ptr(point_list), _ptr(image_state), _ptr(color), _ptr(depth),
            ctypes.c_int32(dmode), ctypes.c_float(settings.depth_near), ctypes.c_float(settings.depth_far),
            _ptr(bin_scratch), st))
        if tracker is not None:
            tracker.observe(counters)
        if settings.debug:
            torch.cuda.synchronize(device)
    del keep
    return color, ForwardState(geom, radii, point_list, image_state, N if settings.instance_capacity is None else -1, nvis,
                               depth, counters)


class CapacityTracker:
    """Sync-free sizing of the instance buffers for a render loop (training steps, video frames).

    The exact path reads the instance count N back after K1 (one small host wait per view).  A loop whose views change
    slowly can skip that wait: render step i with ``instance_capacity`` = (largest N seen so far) x ``margin``, and
    learn step i's own N -- and whether it overflowed -- from an asynchronous copy of the 16-byte device counters that
    is polled WITHOUT blocking at later steps.

        tracker = CapacityTracker()
        for ...:
            s = tracker.settings(s)                      # sets instance_capacity once a count is known
            out = GaussianRasterizer(s)(...)             # or forward_raw
            tracker.observe(state_or_counters)           # async D2H of the counters, no wait
        tracker.flush(); assert not tracker.overflowed   # one wait at the end (or whenever the results are consumed)

    Until the first count has arrived the settings are returned unchanged (exact path).  An overflow can only be
    detected after the fact: ``overflowed`` lists the steps whose buffers were too small (their images / gradients are
    incomplete and should be redone -- with the grown capacity the retry succeeds).

    ``freeze()`` pins the capacities for CUDA-graph capture (graph.GraphedAutogradStep): no polling, no asynchronous
    copies; the tracker only remembers the device counters of the captured call, and ``frozen_overflowed()`` reads them
    (one sync) whenever the caller wants to know whether the last replay fitted."""

    def __init__(self, margin: float = 1.25, min_capacity: int = 4096) -> None:
        self.margin, self.min_capacity = float(margin), int(min_capacity)
        self.capacity: Optional[int] = None
        self.pair_capacity: Optional[int] = None   # batched path: (view, Gaussian) pairs
        self.max_seen = 0
        self.max_pairs = 0
        self.overflowed: list = []
        self._pending: list = []
        self._step = 0
        self.frozen = False
        self._static_counters = None

    def freeze(self) -> None:
        self.flush()
        if self.capacity is None:
            raise RuntimeError("CapacityTracker.freeze: no count has been observed yet (run the call at least once)")
        self.frozen = True

    def unfreeze(self) -> None:
        self.frozen, self._static_counters = False, None

    def frozen_overflowed(self) -> bool:
        """Did the last captured / replayed call need more than the frozen capacities?  Synchronises."""
        c = self._static_counters
        return c is not None and bool(int(c[1].item()) & 3)

    def settings(self, s: GaussianRasterizationSettings, pairs: bool = False) -> GaussianRasterizationSettings:
        if not self.frozen:
            self.poll()
        if self.capacity is None:
            return s
        if pairs:
            return s._replace(instance_capacity=self.capacity, pair_capacity=self.pair_capacity)
        return s._replace(instance_capacity=self.capacity)

    def observe(self, counters) -> None:
        """``counters``: the device int32[4] of a forward call (ForwardState.counters) or the state itself."""
        c = getattr(counters, "counters", counters)
        if self.frozen:
            self._static_counters = c
            return
        host = torch.empty(4, dtype=torch.int32).pin_memory()
        host.copy_(c, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(c.device))
        self._pending.append((self._step, host, ev))
        self._step += 1

    def _take(self, step, host) -> None:
        n, ovf, npairs = int(host[0].item()) & 0xFFFFFFFF, int(host[1].item()), int(host[2].item()) & 0xFFFFFFFF
        self.max_seen, self.max_pairs = max(self.max_seen, n), max(self.max_pairs, npairs)
        want = max(self.min_capacity, int(self.max_seen * self.margin) + 1)
        if self.capacity is None or want > self.capacity:
            self.capacity = want
        want = max(self.min_capacity, int(self.max_pairs * self.margin) + 1)
        if self.pair_capacity is None or want > self.pair_capacity:
            self.pair_capacity = want
        if ovf & 3:
            self.overflowed.append(step)

    def poll(self) -> None:
        while self._pending and self._pending[0][2].query():
            step, host, _ = self._pending.pop(0)
            self._take(step, host)

    def flush(self) -> None:
        for step, host, ev in self._pending:
            ev.synchronize()
            self._take(step, host)
        self._pending = []


def overflowed(state: ForwardState) -> bool:
    """True if a forward call made with ``instance_capacity`` needed more instances than it was given (its image and
    gradients are then incomplete).  Synchronises on the device."""
    return state.counters is not None and bool(int(state.counters[1].item()) != 0)


def instances_needed(state: ForwardState) -> int:
    """Number of (tile, Gaussian) instances the view needed (synchronises); use it to pick ``instance_capacity``."""
    return int(state.counters[0].item()) & 0xFFFFFFFF


def _depth_args(settings: GaussianRasterizationSettings, grad_depth: Optional[Tensor], device):
    """(pointer, mode, near, far) of the fused depth channel's gradient for the backward entry points."""
    if grad_depth is None or settings.depth_mode is None:
        return None, ctypes.c_int32(0), ctypes.c_float(0.0), ctypes.c_float(0.0), None
    g = _f32c(grad_depth, device)
    return (_ptr(g), ctypes.c_int32(_lib.DEPTH_MODES[settings.depth_mode]), ctypes.c_float(settings.depth_near),
            ctypes.c_float(settings.depth_far), g)


def backward_raw(settings: GaussianRasterizationSettings, means3D: Tensor, cov6: Tensor, opacities: Tensor,
                 shs: Optional[Tensor], colors: Optional[Tensor], state: ForwardState, grad_color: Tensor,
                 grad_depth: Optional[Tensor] = None):
    """Run the backward pass through the C-ABI.  Returns a dict of gradient tensors.  ``grad_depth`` [H,W]: gradient
    w.r.t. the fused depth channel (settings.depth_mode)."""
    lib = _lib.load()
    device = means3D.device
    P = means3D.shape[0]
    M = (shs.shape[2] if settings.sh_layout else shs.shape[1]) if shs is not None else 0
    with torch.cuda.device(device):
        view, keep = _make_view(settings, P, M, device)
        f32 = dict(dtype=torch.float32, device=device)
        g_means = torch.empty((P, 3), **f32)
        g_means2D = torch.empty((P, 3), **f32)
        g_cov = torch.empty_like(cov6)
        g_op = torch.empty((P, 1), **f32)
        g_sh = torch.empty_like(shs) if shs is not None else None
        g_col = torch.empty((P, 3), **f32) if colors is not None else None
        scratch = torch.empty(lib.s360_backward_scratch_bytes(P), dtype=torch.uint8, device=device)
        grad_color = _f32c(grad_color, device)
        gd_ptr, gd_mode, gd_near, gd_far, gd_keep = _depth_args(settings, grad_depth, device)
        _lib.check(lib.s360_backward(
            ctypes.byref(view), _ptr(means3D), _ptr(cov6), _ptr(opacities), _ptr(shs), _ptr(colors),
            _ptr(state.geom), _ptr(state.radii), _ptr(state.point_list), _ptr(state.image_state),
            _ptr(grad_color), gd_ptr, gd_mode, gd_near, gd_far,
            _ptr(g_means), _ptr(g_means2D), _ptr(g_cov), _ptr(g_op), _ptr(g_sh), _ptr(g_col),
            _ptr(scratch), _stream_ptr()))
        del gd_keep
        if settings.debug:
            torch.cuda.synchronize(device)
    del keep
    return dict(means3D=g_means, means2D=g_means2D, cov3D=g_cov, opacities=g_op, shs=g_sh, colors=g_col)


# ------------------------------------------------------------------------------------------------
# batched multi-view path: V views of the same Gaussians in one pass
class MultiForwardState(NamedTuple):
    geom: Tensor
    point_list: Tensor
    image_state: Tensor
    views: int
    pair_capacity: int
    num_rendered: int          # -1 when instance_capacity was given (nothing read back)
    num_pairs: int             # (view, Gaussian) pairs that touch a tile; -1 when not read back
    radii: Optional[Tensor] = None     # [V,P] int32 when requested
    depth: Optional[Tensor] = None     # [V,H,W] fused depth channel
    counters: Optional[Tensor] = None  # device int32[4]: num_rendered, overflow bits, pairs needed, -


def forward_views_raw(settings: GaussianRasterizationSettings, means3D: Tensor, cov: Tensor, opacities: Tensor,
                      shs: Optional[Tensor], colors: Optional[Tensor], want_radii: bool = False,
                      _pair_capacity_override: Optional[int] = None):
    """All views of ``settings`` (viewmatrix [V,4,4], projmatrix [V,4,4], campos [V,3]; everything else shared) in one
    pass through the batched C-ABI.  Returns (color [V,3,H,W], MultiForwardState)."""
    lib = _lib.load()
    device = means3D.device
    if device.type != "cuda":
        raise RuntimeError("splatter360_b200 rasterizer needs CUDA tensors (there is no CPU path)")
    V = int(torch.as_tensor(settings.viewmatrix).numel() // 16)
    if not 1 <= V <= _lib.MAX_VIEWS:
        raise ValueError(f"the batched path takes 1..{_lib.MAX_VIEWS} views per pass, got {V}")
    P = means3D.shape[0]
    M = (shs.shape[2] if settings.sh_layout else shs.shape[1]) if shs is not None else 0
    H, W = int(settings.image_height), int(settings.image_width)
    tracker = settings.capacity_tracker
    if tracker is not None and settings.instance_capacity is None and settings.pair_capacity is None:
        settings = tracker.settings(settings, pairs=True)
    if settings.pair_capacity is not None:
        pcap = min(int(settings.pair_capacity), V * P)
    elif _pair_capacity_override is not None:
        pcap = int(_pair_capacity_override)
    elif settings.projection == "erp" or settings.instance_capacity is not None:
        pcap = V * P   # erp: every Gaussian is in every view; the sync-free mode cannot retry: worst case
    else:
        # pinhole views look in different directions (cube faces: ~1.05 pairs per Gaussian): start with 2 per Gaussian,
        # the count read back below triggers one exact retry when that was too small
        pcap = min(V * P, 2 * P + 4096)
    pcap = max(pcap, 1)
    with torch.cuda.device(device):
        view, keep = _make_view(settings, P, M, device, views=V)
        u8 = dict(dtype=torch.uint8, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        geom = torch.empty(lib.s360_multi_geom_bytes(P, pcap), **u8)
        pre_scratch = torch.empty(lib.s360_multi_preprocess_scratch_bytes(P, pcap), **u8)
        radii = torch.empty((V, P), **i32) if want_radii else None
        depth_order = torch.empty(pcap, **i32)
        offsets = torch.empty(pcap, **i32)
        counters = torch.empty(4, **i32)
        st = _stream_ptr()
        head = (ctypes.byref(view), ctypes.c_int32(V), ctypes.c_int64(pcap))
        _lib.check(lib.s360_multi_forward_project(
            *head, _ptr(means3D), _ptr(cov), _ptr(opacities), _ptr(shs), _ptr(colors),
            _ptr(geom), _ptr(radii), _ptr(counters), _ptr(pre_scratch), st))
        if settings.instance_capacity is not None:
            _lib.check(lib.s360_multi_forward_order(
                *head, _ptr(geom), _ptr(depth_order), _ptr(offsets), _ptr(counters), _ptr(pre_scratch), st))
            N, npairs = int(settings.instance_capacity), -1
        else:
            host_counts, side, ready, lock = _count_reader(device)
            with lock:
                ready.record(torch.cuda.current_stream())
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    host_counts.copy_(counters, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(side)
                counters.record_stream(side)
                _lib.check(lib.s360_multi_forward_order(
                    *head, _ptr(geom), _ptr(depth_order), _ptr(offsets), _ptr(counters), _ptr(pre_scratch), st))
                done.synchronize()
                N, npairs = int(host_counts[0].item()) & 0xFFFFFFFF, int(host_counts[2].item()) & 0xFFFFFFFF
            if npairs > pcap:
                if settings.pair_capacity is not None:
                    raise RuntimeError(f"pair_capacity {pcap} too small: this batch needs {npairs} (view, Gaussian) pairs")
                del geom, pre_scratch, depth_order, offsets
                return forward_views_raw(settings, means3D, cov, opacities, shs, colors, want_radii, npairs)
        cap = max(N, 1)
        point_list = torch.empty(cap, **i32)
        bin_scratch = torch.empty(lib.s360_multi_binning_scratch_bytes(pcap, cap, V, H, W), **u8)
        image_state = torch.empty(lib.s360_multi_image_bytes(V, H, W), **u8)
        color = torch.empty((V, 3, H, W), dtype=torch.float32, device=device)
        depth = None
        dmode = 0
        if settings.depth_mode is not None:
            if settings.depth_mode not in _lib.DEPTH_MODES:
                raise ValueError(f"unknown depth_mode {settings.depth_mode!r}")
            dmode = _lib.DEPTH_MODES[settings.depth_mode]
            depth = torch.empty((V, H, W), dtype=torch.float32, device=device)
        _lib.check(lib.s360_multi_forward_render(
            *head, _ptr(geom), _ptr(depth_order), _ptr(offsets), _ptr(counters),
            ctypes.c_int64(N), _ptr(point_list), _ptr(image_state), _ptr(color), _ptr(depth),
            ctypes.c_int32(dmode), ctypes.c_float(settings.depth_near), ctypes.c_float(settings.depth_far),
            _ptr(bin_scratch), st))
        if tracker is not None:
            tracker.observe(counters)
        if settings.debug:
            torch.cuda.synchronize(device)
    del keep
    return color, MultiForwardState(geom, point_list, image_state, V, pcap,
                                    N if settings.instance_capacity is None else -1, npairs, radii, depth, counters)


def backward_views_raw(settings: GaussianRasterizationSettings, means3D: Tensor, cov: Tensor, opacities: Tensor,
                       shs: Optional[Tensor], colors: Optional[Tensor], state: MultiForwardState, grad_color: Tensor,
                       grad_depth: Optional[Tensor] = None):
    """Backward of ``forward_views_raw``: grad_color [V,3,H,W] (and grad_depth [V,H,W] of the fused depth channel)
    -> gradients summed over the views."""
    lib = _lib.load()
    device = means3D.device
    P = means3D.shape[0]
    M = (shs.shape[2] if settings.sh_layout else shs.shape[1]) if shs is not None else 0
    V = state.views
    with torch.cuda.device(device):
        view, keep = _make_view(settings, P, M, device, views=V)
        f32 = dict(dtype=torch.float32, device=device)
        g_means = torch.empty((P, 3), **f32)
        g_cov = torch.empty_like(cov)
        g_op = torch.empty((P, 1), **f32)
        g_sh = torch.empty_like(shs) if shs is not None else None
        g_col = torch.empty((P, 3), **f32) if colors is not None else None
        scratch = torch.empty(lib.s360_multi_backward_scratch_bytes(state.pair_capacity), dtype=torch.uint8, device=device)
        grad_color = _f32c(grad_color, device)
        gd_ptr, gd_mode, gd_near, gd_far, gd_keep = _depth_args(settings, grad_depth, device)
        _lib.check(lib.s360_multi_backward(
            ctypes.byref(view), ctypes.c_int32(V), ctypes.c_int64(state.pair_capacity),
            _ptr(means3D), _ptr(cov), _ptr(opacities), _ptr(shs), _ptr(colors),
            _ptr(state.geom), _ptr(state.point_list), _ptr(state.image_state), _ptr(grad_color),
            gd_ptr, gd_mode, gd_near, gd_far,
            _ptr(g_means), _ptr(g_cov), _ptr(g_op), _ptr(g_sh), _ptr(g_col), _ptr(scratch), _stream_ptr()))
        del gd_keep
        if settings.debug:
            torch.cuda.synchronize(device)
    del keep
    return dict(means3D=g_means, cov3D=g_cov, opacities=g_op, shs=g_sh, colors=g_col)


class _RasterizeViews(torch.autograd.Function):
    """V views in one pass.  Inputs as ``_RasterizeGaussians`` minus means2D/scales/rotations; outputs
    color [V,3,H,W] (+ depth [V,H,W] when settings.depth_mode is set)."""

    @staticmethod
    def forward(ctx, means3D, sh, colors_precomp, opacities, cov3Ds_precomp, raster_settings):
        device = means3D.device
        means3D_c = _f32c(means3D, device)
        cov = _f32c(cov3Ds_precomp, device)
        op = _f32c(opacities, device).reshape(-1)
        shs_c = _f32c(sh, device) if sh is not None and sh.numel() else None
        col_c = _f32c(colors_precomp, device) if colors_precomp is not None and colors_precomp.numel() else None
        P = means3D_c.shape[0]
        if cov.shape != ((P, 3, 3) if raster_settings.cov_layout else (P, 6)) or op.shape[0] != P:
            raise ValueError("cov3D_precomp must be [P,6] ([P,3,3] with cov_layout=1) and opacities [P] or [P,1]")
        if (shs_c is None) == (col_c is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        color, state = forward_views_raw(raster_settings, means3D_c, cov, op, shs_c, col_c)
        ctx.raster_settings = raster_settings
        ctx.state = state
        ctx.has_sh = shs_c is not None
        ctx.op_shape = opacities.shape
        ctx.save_for_backward(means3D_c, cov, op, shs_c if shs_c is not None else col_c)
        ctx.out_shape = color.shape
        if state.depth is not None:
            return color, state.depth
        return color

    @staticmethod
    def backward(ctx, grad_out_color, grad_depth=None):
        means3D, cov, op, feat = ctx.saved_tensors
        shs = feat if ctx.has_sh else None
        col = None if ctx.has_sh else feat
        if grad_out_color is None:   # only the depth channel entered the loss
            grad_out_color = torch.zeros(ctx.out_shape, dtype=torch.float32, device=means3D.device)
        g = backward_views_raw(ctx.raster_settings, means3D, cov, op, shs, col, ctx.state, grad_out_color, grad_depth)
        return (g["means3D"], g["shs"], g["colors"], g["opacities"].reshape(ctx.op_shape), g["cov3D"], None)


def rasterize_views(means3D, opacities, cov3D_precomp, raster_settings, shs=None, colors_precomp=None):
    """Render every camera of ``raster_settings`` (viewmatrix [V,4,4], projmatrix [V,4,4], campos [V,3]) in ONE pass:
    color [V,3,H,W] (and depth [V,H,W] with ``depth_mode``).  Same values as V ``GaussianRasterizer`` calls; the
    gradients are their sum.  ``means2D`` screen-space gradients are not produced."""
    return _RasterizeViews.apply(means3D, shs, colors_precomp, opacities, cov3D_precomp, raster_settings)


def _cpu_copy(args):
    """Upstream's ``cpu_deep_copy_tuple``: what a debug snapshot holds."""
    return tuple(a.detach().cpu().clone() if isinstance(a, Tensor) else (_cpu_copy(a) if isinstance(a, tuple) else a)
                 for a in args)


def _debug_guard(debug: bool, which: str, args, fn):
    """Upstream's debug convention: with ``settings.debug`` an exception in the native call dumps the
    CPU copies of its arguments to ``snapshot_fw.dump`` / ``snapshot_bw.dump`` and is re-raised."""
    if not debug:
        return fn()
    cpu_args = _cpu_copy(args)
    try:
        return fn()
    except Exception as ex:
        torch.save(cpu_args, f"snapshot_{which}.dump")
        print(f"\nAn error occured in {'forward' if which == 'fw' else 'backward'}. "
              f"Please forward snapshot_{which}.dump for debugging.")
        raise ex


class _RasterizeGaussians(torch.autograd.Function):
    """Same forward/backward signature as upstream's autograd.Function."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        device = means3D.device
        means3D_c = _f32c(means3D, device)
        cov6 = _f32c(cov3Ds_precomp, device)
        op = _f32c(opacities, device).reshape(-1)
        shs_c = _f32c(sh, device) if sh.numel() else None
        col_c = _f32c(colors_precomp, device) if colors_precomp.numel() else None
        P = means3D_c.shape[0]
        if cov6.shape != ((P, 3, 3) if raster_settings.cov_layout else (P, 6)) or op.shape[0] != P:
            raise ValueError("cov3D_precomp must be [P,6] ([P,3,3] with cov_layout=1) and opacities [P,1]")
        if shs_c is not None and (shs_c.dim() != 3 or shs_c.shape[0] != P or shs_c.shape[1 if raster_settings.sh_layout else 2] != 3):
            raise ValueError("shs must be [P,M,3] ([P,3,M] with sh_layout=1)")
        if col_c is not None and col_c.shape != (P, 3):
            raise ValueError("colors_precomp must be [P,3]")
        color, state = _debug_guard(
            bool(raster_settings.debug), "fw", (means3D_c, cov6, op, shs_c, col_c, tuple(raster_settings)),
            lambda: forward_raw(raster_settings, means3D_c, cov6, op, shs_c, col_c))
        ctx.raster_settings = raster_settings
        ctx.state = state
        ctx.has_sh = shs_c is not None
        ctx.save_for_backward(means3D_c, cov6, op, shs_c if shs_c is not None else col_c)
        ctx.mark_non_differentiable(state.radii)
        ctx.out_shape = color.shape
        if state.depth is not None:
            return color, state.radii, state.depth
        return color, state.radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, grad_depth=None):
        means3D, cov6, op, feat = ctx.saved_tensors
        shs = feat if ctx.has_sh else None
        col = None if ctx.has_sh else feat
        if grad_out_color is None:   # only the depth channel entered the loss
            grad_out_color = torch.zeros(ctx.out_shape, dtype=torch.float32, device=means3D.device)
        g = _debug_guard(
            bool(ctx.raster_settings.debug), "bw", (means3D, cov6, op, feat, grad_out_color, grad_depth, tuple(ctx.raster_settings)),
            lambda: backward_raw(ctx.raster_settings, means3D, cov6, op, shs, col, ctx.state, grad_out_color, grad_depth))
        return (g["means3D"], g["means2D"], g["shs"], g["colors"], g["opacities"], None, None, g["cov3D"], None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: Tensor) -> Tensor:
        """Boolean mask of Gaussians in front of the near-cull distance (upstream ``mark_visible``)."""
        lib = _lib.load()
        s = self.raster_settings
        with torch.no_grad():
            pos = _f32c(positions, positions.device)
            if pos.device.type != "cuda":
                raise RuntimeError("markVisible needs CUDA tensors")
            P = pos.shape[0]
            with torch.cuda.device(pos.device):
                view, keep = _make_view(s, P, 0, pos.device)
                out = torch.empty(P, dtype=torch.uint8, device=pos.device)
                _lib.check(lib.s360_mark_visible(ctypes.byref(view), _ptr(pos), _ptr(out), _stream_ptr()))
            del keep
            return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        s = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
        if shs is None:
            shs = empty
        if colors_precomp is None:
            colors_precomp = empty
        if cov3D_precomp is None:
            cov3D_precomp = build_covariance_6(scales, rotations, float(s.scale_modifier))
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, empty, empty, cov3D_precomp, s)
