"""Fused Gaussian adapter: same call contract as the reference's ``GaussianAdapterERP``
(/root/reference/src/model/encoder/common/gaussian_adapter_erp.py:34-119), one CUDA kernel forward and one backward
instead of ~25 torch ops:

    adapter = GaussianAdapterERP(GaussianAdapterERPCfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, sh_degree=4))
    g = adapter.forward("hm3d", extrinsics[b,v,1,1,1,4,4], depths[b,v,r,1,1], opacities[b,v,r,1,1],
                        raw_gaussians[b,v,r,1,1,7+3*d_sh], (h, w))          # r = h*w, one Gaussian per context pixel
    g.means [b,v,r,1,1,3]  g.covariances [...,3,3]  g.harmonics [...,3,d_sh]  g.opacities  g.scales  g.rotations

Per view the only host-side work is the SH rotation matrix of the camera-to-world rotation (Wigner D^0..D^4, 165 floats),
built here in torch from e3nn's published construction (Y-X-Y Euler angles, real-basis generators; e3nn itself is an absent
dependency -- see oracle/e3nn_wigner.py for the pinned-by-properties restatement this mirrors).

Like the reference, the means carry no gradient (its unprojection runs under ``torch.no_grad()``, sphere_projection.py:14);
``means_grad=True`` enables d mean / d depth.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from functools import lru_cache

import torch
from torch import Tensor, nn

from . import _lib


@dataclass
class Gaussians:
    """/root/reference/src/model/encoder/common/gaussian_adapter.py:13-20."""
    means: Tensor
    covariances: Tensor
    scales: Tensor
    rotations: Tensor
    harmonics: Tensor
    opacities: Tensor


@dataclass
class GaussianAdapterERPCfg:
    gaussian_scale_min: float
    gaussian_scale_max: float
    sh_degree: int


# ---- SH rotation matrices (host side, per view) ------------------------------------------------------------------
def _su2_generators(j: int) -> Tensor:
    m = torch.arange(-j, j, dtype=torch.float64)
    raising = torch.diag(-torch.sqrt(j * (j + 1) - m * (m + 1)), diagonal=-1).to(torch.complex128)
    m = torch.arange(-j + 1, j + 1, dtype=torch.float64)
    lowering = torch.diag(torch.sqrt(j * (j + 1) - m * (m - 1)), diagonal=1).to(torch.complex128)
    m = torch.arange(-j, j + 1, dtype=torch.float64)
    return torch.stack([0.5 * (raising + lowering), torch.diag(1j * m.to(torch.complex128)), -0.5j * (raising - lowering)])


@lru_cache(maxsize=None)
def _so3_generators(l: int) -> Tensor:
    """Real-basis generators of the degree-l representation (x and y axes), e3nn convention."""
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / 2 ** 0.5
        q[l + m, l - abs(m)] = -1j / 2 ** 0.5
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / 2 ** 0.5
        q[l + m, l - abs(m)] = 1j * (-1) ** m / 2 ** 0.5
    q = (-1j) ** l * q
    X = torch.conj(q.T) @ _su2_generators(l) @ q
    return torch.real(X)


def sh_rotation_blocks(rotations: Tensor, sh_degree: int, mask: Tensor) -> Tensor:
    """[n,3,3] rotations -> [n,165] float32: row-major D^0 | D^1 | ... | D^4 (zero beyond sh_degree) with column j of band l
    scaled by mask[l*l + j] -- what ``rotate_sh(sh * mask, R)`` (sh_rotation.py:10-30) multiplies the raw coefficients with."""
    R = rotations.detach().double().cpu()
    n = R.shape[0]
    y = R @ torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    y = torch.nn.functional.normalize(y, dim=-1).clamp(-1, 1)
    alpha, beta = torch.atan2(y[:, 0], y[:, 2]), torch.acos(y[:, 1])

    def ry(a):
        c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
        return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)

    def rx(a):
        c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
        return torch.stack([torch.stack([o, z, z], -1), torch.stack([z, c, -s], -1), torch.stack([z, s, c], -1)], -2)

    Rr = (ry(alpha) @ rx(beta)).transpose(-1, -2) @ R
    gamma = torch.atan2(Rr[:, 0, 2], Rr[:, 0, 0])
    out = torch.zeros(n, 165, dtype=torch.float64)
    off = 0
    mask = mask.detach().double().cpu()
    for l in range(5):
        k = 2 * l + 1
        if l <= sh_degree:
            X = _so3_generators(l)
            ex = lambda a, G: torch.matrix_exp((a % (2 * math.pi))[:, None, None] * G)
            D = ex(alpha, X[1]) @ ex(beta, X[0]) @ ex(gamma, X[1])
            D = D * mask[l * l:(l + 1) ** 2][None, None, :]
            out[:, off:off + k * k] = D.reshape(n, k * k)
        off += k * k
    return out.float()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Adapter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, depth, pose, rot, H, W, sh_degree, smin, smax, means_grad):
        lib = _lib.load()
        if raw.device.type != "cuda":
            raise RuntimeError("splatter360_b200 adapter needs CUDA tensors (there is no CPU path)")
        G, C = raw.shape
        views = G // (H * W)
        d_sh = (sh_degree + 1) ** 2
        f32 = dict(dtype=torch.float32, device=raw.device)
        means = torch.empty((G, 3), **f32); cov = torch.empty((G, 3, 3), **f32); sh = torch.empty((G, 3, d_sh), **f32)
        scales = torch.empty((G, 3), **f32); quat = torch.empty((G, 4), **f32)
        with torch.cuda.device(raw.device):
            _lib.check(lib.s360_adapter_forward(views, H, W, sh_degree, smin, smax, _p(raw), _p(depth), _p(pose), _p(rot),
                                                _p(means), _p(cov), _p(sh), _p(scales), _p(quat),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.save_for_backward(raw, depth, pose, rot)
        ctx.cfg = (views, H, W, sh_degree, smin, smax, means_grad)
        ctx.mark_non_differentiable(scales, quat)
        return means, cov, sh, scales, quat

    @staticmethod
    def backward(ctx, g_means, g_cov, g_sh, _gs, _gq):
        lib = _lib.load()
        raw, depth, pose, rot = ctx.saved_tensors
        views, H, W, sh_degree, smin, smax, means_grad = ctx.cfg
        c = lambda t: None if t is None else t.contiguous().float()
        g_means, g_cov, g_sh = c(g_means), c(g_cov), c(g_sh)
        d_raw = torch.empty_like(raw); d_depth = torch.empty_like(depth)
        with torch.cuda.device(raw.device):
            _lib.check(lib.s360_adapter_backward(views, H, W, sh_degree, smin, smax, int(means_grad), _p(raw), _p(depth), _p(pose),
                                                 _p(rot), _p(g_means), _p(g_cov), _p(g_sh), _p(d_raw), _p(d_depth),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return d_raw, d_depth, None, None, None, None, None, None, None, None


class GaussianAdapterERP(nn.Module):
    def __init__(self, cfg: GaussianAdapterERPCfg, means_grad: bool = False) -> None:
        super().__init__()
        self.cfg = cfg
        self.means_grad = means_grad
        self.register_buffer("sh_mask", torch.ones((self.d_sh,), dtype=torch.float32), persistent=False)
        for degree in range(1, cfg.sh_degree + 1):
            self.sh_mask[degree ** 2:(degree + 1) ** 2] = 0.1 * 0.25 ** degree

    @property
    def d_sh(self) -> int:
        return (self.cfg.sh_degree + 1) ** 2

    @property
    def d_in(self) -> int:
        return 7 + 3 * self.d_sh

    def forward(self, dataset_name: str, extrinsics: Tensor, depths: Tensor, opacities: Tensor, raw_gaussians: Tensor,
                image_shape: tuple, eps: float = 1e-8) -> Gaussians:
        """Argument meaning as the reference (gaussian_adapter_erp.py:49-60); ``dataset_name`` selects the ERP pixel
        convention -- only the reference's hm3d / replica convention exists here."""
        if self.cfg.sh_degree > 4:
            raise ValueError("sh_degree <= 4")
        if eps != 1e-8:
            raise ValueError("eps is fixed at the reference's default 1e-8")
        h, w = image_shape
        b, v = depths.shape[:2]
        r = h * w
        if depths.shape[2] != r or raw_gaussians.shape[-1] != self.d_in or tuple(raw_gaussians.shape[3:5]) != (1, 1):
            raise ValueError("expected one Gaussian per context pixel: depths [b,v,h*w,1,1], raw [b,v,h*w,1,1,7+3*d_sh]")
        dev = raw_gaussians.device
        ext = extrinsics.reshape(b * v, 4, 4).float()
        pose = torch.cat([ext[:, :3, :3].reshape(b * v, 9), ext[:, :3, 3]], dim=-1).contiguous()
        rot = sh_rotation_blocks(ext[:, :3, :3], self.cfg.sh_degree, self.sh_mask).to(dev)
        raw = raw_gaussians.reshape(b * v * r, self.d_in).float().contiguous()
        dep = depths.reshape(b * v * r).float().contiguous()
        means, cov, sh, scales, quat = _Adapter.apply(raw, dep, pose, rot, h, w, self.cfg.sh_degree,
                                                      float(self.cfg.gaussian_scale_min), float(self.cfg.gaussian_scale_max),
                                                      self.means_grad)
        sh5 = (b, v, r, 1, 1)
        return Gaussians(means=means.reshape(*sh5, 3), covariances=cov.reshape(*sh5, 3, 3),
                         harmonics=sh.reshape(*sh5, 3, self.d_sh), opacities=opacities,
                         scales=scales.reshape(*sh5, 3), rotations=quat.reshape(*sh5, 4))
