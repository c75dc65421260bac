"""Cube-face conventions of the reference and the cube -> equirectangular stitch, for code that still wants the
reference's six-face outputs and for cross-checking the native ``erp`` mode against them.

* ``cube_face_extrinsics`` : sphere camera-to-world -> the six OpenCV face poses in dataset order [top, front, left, back,
  right, bottom] = py360 [U B L F R D]  (/root/reference/preprocess/convert_cubemaps_mp.py:151-193: c2w @ R_k, then the
  y and z axes are negated to go from the habitat to the OpenCV convention)
* ``change_order``         : [U B L F R D] -> [F R B L U D] with U and D flipped in both image axes
  (/root/reference/src/model/model_wrapper_erp.py:135-158)
* ``Cube2Equirec``         : the stitch of /root/reference/src/geometry/layers.py:41-116 (face selection per ERP pixel,
  per-face tangent-plane coordinates, bilinear lookup with border clamp); pinned against the reference's own output in
  tests/test_golden.py
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor, nn


def _rx(deg: float) -> Tensor:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[1.0, 0, 0], [0, c, -s], [0, s, c]])


def _ry(deg: float) -> Tensor:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0, s], [0, 1.0, 0], [-s, 0, c]])


def cube_face_extrinsics(extrinsics_sphere: Tensor) -> Tensor:
    """[..., 4, 4] sphere pose -> [..., 6, 4, 4] OpenCV camera-to-world of the faces [top, front, left, back, right, bottom]."""
    rots = torch.stack([_rx(90), torch.eye(3), _ry(-90), _ry(-180), _ry(-270), _rx(-90)]).to(extrinsics_sphere)
    out = extrinsics_sphere[..., None, :, :].repeat(*([1] * (extrinsics_sphere.dim() - 2)), 6, 1, 1)
    out[..., :3, :3] = extrinsics_sphere[..., None, :3, :3] @ rots
    out[..., :, 1] = -out[..., :, 1]
    out[..., :, 2] = -out[..., :, 2]
    return out


def change_order(cubes: Tensor) -> Tensor:
    """[6, ...] faces in dataset order [U B L F R D] -> [F R B L U D]; U and D are flipped along both image axes."""
    up = torch.flip(cubes[0], dims=[-1, -2])
    down = torch.flip(cubes[5], dims=[-1, -2])
    return torch.stack([cubes[3], cubes[4], cubes[1], cubes[2], up, down])


def depth_to_distance_factor(face_w: int, fx: float, fy: float, cx: float, cy: float) -> Tensor:
    """[f, f] factor distance / z-depth of every texel, literal to the reference's ``depth_to_distance_map_batch``
    (/root/reference/src/geometry/z_depth_to_distance.py:4-34): integer pixel coordinates, and its default-indexed
    ``torch.meshgrid(arange(width), arange(height))`` makes "u" (cx, fx) run along the rows.  Torch formulation: the checker
    of the fused kernel option (``Cube2Equirec.from_faces(depth_to_distance=...)``)."""
    r = torch.arange(face_w, dtype=torch.float32)[:, None]
    c = torch.arange(face_w, dtype=torch.float32)[None, :]
    return torch.sqrt(((r - cx) / fx) ** 2 + ((c - cy) / fy) ** 2 + 1.0)


class Cube2Equirec(nn.Module):
    """faces [b, c, f, 6 f] laid out side by side in the order [F R B L U D] -> panorama [b, c, H, W]."""

    def __init__(self, face_w: int, equ_h: int, equ_w: int) -> None:
        super().__init__()
        self.face_w, self.equ_h, self.equ_w = face_w, equ_h, equ_w
        H, W = equ_h, equ_w
        xs = torch.arange(W)
        # side face of every column: the four quadrants of longitude, front centred on the image
        side = ((xs - 3 * W // 8) % W) // (W // 4)
        tp = side[None, :].repeat(H, 1)
        # rows above the boundary latitude atan(cos(lon')) belong to the top face, mirrored rows to the bottom face
        q = torch.linspace(-math.pi, math.pi, W // 4, dtype=torch.float64) / 4
        first_side_row = H // 2 - torch.round(torch.atan(torch.cos(q)) * H / math.pi).long()
        rows = torch.arange(H)[:, None]
        up_q = rows < first_side_row[None, :]                                 # [H, W/4]
        up = torch.roll(up_q.repeat(1, 4), 3 * W // 8, dims=1)
        tp = torch.where(up, torch.full_like(tp, 4), tp)
        tp = torch.where(torch.flip(up, dims=[0]), torch.full_like(tp, 5), tp)
        lon = ((torch.arange(W, dtype=torch.float32) + 0.5) / W - 0.5) * 2 * math.pi
        lat = -((torch.arange(H, dtype=torch.float32) + 0.5) / H - 0.5) * math.pi
        lat, lon = torch.meshgrid(lat, lon, indexing="ij")
        u = torch.zeros(H, W)
        v = torch.zeros(H, W)
        for i in range(4):
            m = tp == i
            rel = lon[m] - math.pi * i / 2
            u[m] = 0.5 * torch.tan(rel)
            v[m] = -0.5 * torch.tan(lat[m]) / torch.cos(rel)
        m = tp == 4
        c = 0.5 * torch.tan(math.pi / 2 - lat[m])
        u[m] = c * torch.sin(lon[m])
        v[m] = c * torch.cos(lon[m])
        m = tp == 5
        c = 0.5 * torch.tan(math.pi / 2 - lat[m].abs())
        u[m] = c * torch.sin(lon[m])
        v[m] = -c * torch.cos(lon[m])
        grid = torch.stack([u.clamp(-0.5, 0.5) * 2, v.clamp(-0.5, 0.5) * 2, tp.float() / 2.5 - 1], dim=-1)
        self.register_buffer("sample_grid", grid.view(1, 1, H, W, 3), persistent=False)

    def forward_reference(self, cube_feat: Tensor) -> Tensor:
        """The reference's own formulation (5-D ``F.grid_sample``, layers.py:108-116), pinned against the reference's
        output in tests/test_golden.py; the checker of the CUDA kernel, not the product path."""
        bs, ch, h, w = cube_feat.shape
        assert h == self.face_w and w == 6 * self.face_w
        faces = cube_feat.view(bs, ch, h, 6, self.face_w).permute(0, 1, 3, 2, 4)      # [b, c, 6, f, f]
        grid = self.sample_grid.expand(bs, -1, -1, -1, -1)
        return F.grid_sample(faces, grid, padding_mode="border", align_corners=True).squeeze(2)

    def forward(self, cube_feat: Tensor) -> Tensor:
        """Strip [b, c, f, 6f] in the order [F R B L U D] -> panorama [b, c, H, W] (libsplatter360 gather kernel)."""
        bs, ch, h, w = cube_feat.shape
        assert h == self.face_w and w == 6 * self.face_w
        return _Cube2EquirecFn.apply(cube_feat, self.sample_grid, 0, self.face_w, self.equ_h, self.equ_w, None)

    def from_faces(self, faces: Tensor, depth_to_distance=None) -> Tensor:
        """Rasterizer output [b, 6, c, f, f] in the dataset face order [U B L F R D] -> panorama [b, c, H, W]:
        ``change_order`` + strip concatenation + stitch in one kernel.

        ``depth_to_distance=(fx, fy, cx, cy)`` (pixels): the faces hold z-depth and the panorama comes out in radial
        distance -- the reference's ``depth_to_distance_map_batch`` between ``change_order`` and the stitch
        (model_wrapper_erp.py:447-463, z_depth_to_distance.py:4-34), fused into the gather."""
        assert faces.dim() == 5 and faces.shape[1] == 6 and faces.shape[-1] == faces.shape[-2] == self.face_w
        k = None if depth_to_distance is None else tuple(float(x) for x in depth_to_distance)
        return _Cube2EquirecFn.apply(faces, self.sample_grid, 1, self.face_w, self.equ_h, self.equ_w, k)


class _Cube2EquirecFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, faces, grid, layout, face_w, H, W, d2d):
        import ctypes
        from . import _lib
        if faces.device.type != "cuda":
            raise RuntimeError("Cube2Equirec needs CUDA tensors (use forward_reference for the torch formulation)")
        lib = _lib.load()
        faces_c = faces.float().contiguous()
        grid_c = grid.to(faces.device).float().contiguous().view(H, W, 3)
        B = faces_c.shape[0]
        C = faces_c.shape[1] if layout == 0 else faces_c.shape[2]
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=faces.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(faces.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            k = None if d2d is None else (ctypes.c_float * 4)(*d2d)
            _lib.check(lib.s360_cube2equirec_forward(p(faces_c), p(grid_c), layout, B, C, face_w, H, W, k, p(out), st))
        ctx.save_for_backward(grid_c)
        ctx.meta = (layout, face_w, H, W, B, C, faces_c.shape, d2d)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        import ctypes
        from . import _lib
        lib = _lib.load()
        (grid_c,) = ctx.saved_tensors
        layout, face_w, H, W, B, C, shape, d2d = ctx.meta
        g = grad_out.float().contiguous()
        d_faces = torch.empty(shape, dtype=torch.float32, device=g.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(g.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            k = None if d2d is None else (ctypes.c_float * 4)(*d2d)
            _lib.check(lib.s360_cube2equirec_backward(p(g), p(grid_c), layout, B, C, face_w, H, W, k, p(d_faces), st))
        return d_faces, None, None, None, None, None, None
