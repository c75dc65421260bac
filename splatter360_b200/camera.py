"""Camera helpers for the rasterizer boundary.

Mirrors, argument for argument, the camera construction the reference performs before every
rasterizer call:

* ``get_fov``                 <- /root/reference/src/geometry/projection.py:233-247
* ``get_projection_matrix``   <- /root/reference/src/model/decoder/cuda_splatting.py:17-44
* ``pinhole_camera``          <- cuda_splatting.py:63-71, 80-87, 109 (1/near rescale, view/full-proj
                                 matrices in the row-vector convention, campos)
* ``erp_camera``              -- native ERP mode: same view-matrix convention, sphere-camera frame of
                                 /root/reference/src/geometry/utils360.py:93-104,148-153,193-198,250-263
* ``erp_pixel_dirs``          <- utils360.py:100-101 (equi_2_spherical) + :151-153 (spherical_2_cartesian)
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch
from torch import Tensor


def inverse(m: Tensor) -> Tensor:
    """``m.inverse()`` of [..., 4, 4] camera matrices (the reference: cuda_splatting.py:84, :176, :262) without the host round
    trip: torch's ``inverse`` is a batched LU over several launches that reads its status back from the device to raise on
    singular input, which stalls an otherwise asynchronous render loop once per call.  CUDA float32 4x4 matrices that need
    no gradient go through one launch of ``s360_invert4x4`` (also capturable in a CUDA graph); everything else (CPU tensors
    of the argument-capture tests, 3x3 intrinsics, differentiable poses) through ``linalg.inv_ex``, which computes the same
    inverse and leaves the status on the device (camera matrices are never singular)."""
    if (m.is_cuda and m.dtype == torch.float32 and m.shape[-2:] == (4, 4) and m.numel() > 0
            and not (m.requires_grad and torch.is_grad_enabled())):
        import ctypes
        from . import _lib
        lib = _lib.load()
        src = m.detach().contiguous()
        out = torch.empty_like(src)
        with torch.cuda.device(m.device):
            _lib.check(lib.s360_invert4x4(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                          ctypes.c_int64(src.numel() // 16),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out
    return torch.linalg.inv_ex(m).inverse


_EDGE_RAYS: dict = {}


def _edge_rays(device) -> Tensor:
    """[4, 3] image-plane points (left, right, top, bottom) on ``device``, built once: ``torch.tensor(list, device=cuda)`` is a
    blocking pageable-memory upload, and the reference's ``get_fov`` does four of them per call."""
    key = str(device)
    t = _EDGE_RAYS.get(key)
    if t is None:
        t = torch.tensor([[0, 0.5, 1], [1, 0.5, 1], [0.5, 0, 1], [0.5, 1, 1]], dtype=torch.float32).to(device)
        _EDGE_RAYS[key] = t
    return t


def get_fov(intrinsics: Tensor) -> Tensor:
    """Field of view [b,2] (x, y) from normalised intrinsics [b,3,3] via the edge-ray angle."""
    inv = inverse(intrinsics)
    rays = torch.einsum("bij,kj->bki", inv, _edge_rays(intrinsics.device))      # [b, 4, 3]
    rays = rays / rays.norm(dim=-1, keepdim=True)
    fov_x = (rays[:, 0] * rays[:, 1]).sum(dim=-1).acos()
    fov_y = (rays[:, 2] * rays[:, 3]).sum(dim=-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor) -> Tensor:
    """[b,4,4]; x/y -> (-1,1), z -> (0,1), w = z (see cuda_splatting.py:17-44)."""
    tan_x = (0.5 * fov_x).tan()
    tan_y = (0.5 * fov_y).tan()
    top = tan_y * near
    bottom = -top
    right = tan_x * near
    left = -right
    (b,) = near.shape
    m = torch.zeros((b, 4, 4), dtype=torch.float32, device=near.device)
    m[:, 0, 0] = 2 * near / (right - left)
    m[:, 1, 1] = 2 * near / (top - bottom)
    m[:, 0, 2] = (right + left) / (right - left)
    m[:, 1, 2] = (top + bottom) / (top - bottom)
    m[:, 3, 2] = 1
    m[:, 2, 2] = far / (far - near)
    m[:, 2, 3] = -(far * near) / (far - near)
    return m


class Camera(NamedTuple):
    """Per-view rasterizer camera block (all batched on dim 0)."""
    view_matrix: Tensor      # [b,4,4] transposed world->camera (p_view = [x y z 1] @ V)
    full_projection: Tensor  # [b,4,4] view_matrix @ proj^T
    tan_fov_x: Tensor        # [b]
    tan_fov_y: Tensor        # [b]
    campos: Tensor           # [b,3]


def pinhole_camera(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor) -> Camera:
    """cuda_splatting.py:80-87.  ``extrinsics`` is camera-to-world (OpenCV), already rescaled."""
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    proj = get_projection_matrix(near, far, fov_x, fov_y).transpose(1, 2)
    view = inverse(extrinsics).transpose(1, 2)
    return Camera(view, view @ proj, (0.5 * fov_x).tan(), (0.5 * fov_y).tan(), extrinsics[:, :3, 3])


def erp_camera(extrinsics_sphere: Tensor) -> Camera:
    """Native ERP camera: only the view matrix and campos are meaningful; the projection slot
    carries the view matrix (unused by the erp kernels) and tan_fov is 1."""
    view = inverse(extrinsics_sphere).transpose(1, 2)
    b = extrinsics_sphere.shape[0]
    one = torch.ones(b, dtype=torch.float32, device=extrinsics_sphere.device)
    return Camera(view, view.clone(), one, one, extrinsics_sphere[:, :3, 3])


def erp_pixel_dirs(height: int, width: int, device=None) -> Tensor:
    """Unit ray directions [H,W,3] of ERP pixel centres in the sphere-camera frame
    (theta=(0.5-(x+.5)/W)2pi, phi=-((y+.5)/H-.5)pi; dir=(cos phi sin theta, sin phi, cos phi cos theta))."""
    xs = torch.arange(width, dtype=torch.float32, device=device)
    ys = torch.arange(height, dtype=torch.float32, device=device)
    theta = (0.5 - (xs + 0.5) / width) * 2 * math.pi
    phi = -((ys + 0.5) / height - 0.5) * math.pi
    phi, theta = torch.meshgrid(phi, theta, indexing="ij")
    return torch.stack((phi.cos() * theta.sin(), phi.sin(), phi.cos() * theta.cos()), dim=-1)


def erp_project(points_cam: Tensor, height: int, width: int) -> Tensor:
    """Sphere-camera-frame points [...,3] -> continuous ERP pixel coords [...,2] (utils360.py:194-198,262-263)."""
    x, y, z = points_cam.unbind(-1)
    theta = torch.atan2(x, z)
    phi = torch.atan2(y, torch.sqrt(x * x + z * z))
    u = (-theta / (2 * math.pi) + 0.5) * width - 0.5
    v = (-phi / math.pi + 0.5) * height - 0.5
    return torch.stack((u, v), dim=-1)
