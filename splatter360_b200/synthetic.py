"""Synthetic Gaussian scenes with the statistics of splatter360's encoder output.

There is no dataset or checkpoint access, so the benchmark and the parity tests draw inputs with the
same laws the reference's encoder/adapter produce:

* scale = (0.5 + 14.5 sigmoid(n)) * depth / max(w,h)   /root/reference/src/model/encoder/common/gaussian_adapter_erp.py:65-77
* cov = R diag(s^2) R^T, xyzw quaternion                 /root/reference/src/model/encoder/common/gaussians.py:8-44
* SH mask 0.1 * 0.25^degree on higher bands              gaussian_adapter_erp.py:46-47
* one Gaussian per ERP pixel per context view            /root/reference/src/model/encoder/encoder_costvolume.py:490-507
* near 0.1 / far 10                                      /root/reference/config/experiment/hm3d.yaml:46-47
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch
from torch import Tensor

from .camera import erp_pixel_dirs


class Scene(NamedTuple):
    means: Tensor        # [G,3]
    covariances: Tensor  # [G,3,3]
    harmonics: Tensor    # [G,3,d_sh]   (reference layout: xyz, then coefficient)
    opacities: Tensor    # [G]


def quaternion_to_matrix(q: Tensor, eps: float = 1e-8) -> Tensor:
    i, j, k, r = torch.unbind(q, dim=-1)
    two_s = 2 / ((q * q).sum(dim=-1) + eps)
    o = torch.stack(
        (1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
         two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
         two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(*q.shape[:-1], 3, 3)


def _attributes(depth: Tensor, sh_degree: int, ref_width: int, gen: torch.Generator):
    n = depth.shape[0]
    dev = depth.device
    scale = (0.5 + 14.5 * torch.randn(n, 3, generator=gen, device=dev).sigmoid()) * depth[:, None] / ref_width
    q = torch.randn(n, 4, generator=gen, device=dev)
    q = q / (q.norm(dim=-1, keepdim=True) + 1e-8)
    rot = quaternion_to_matrix(q)
    cov = rot @ torch.diag_embed(scale * scale) @ rot.transpose(-1, -2)
    opac = torch.randn(n, generator=gen, device=dev).sigmoid()
    d_sh = (sh_degree + 1) ** 2
    sh = torch.randn(n, 3, d_sh, generator=gen, device=dev)
    rgb = torch.rand(n, 3, generator=gen, device=dev)
    sh[:, :, 0] = (rgb - 0.5) / 0.28209479177387814
    for deg in range(1, sh_degree + 1):
        sh[:, :, deg * deg:(deg + 1) ** 2] *= 0.1 * 0.25 ** deg
    return cov, sh, opac


def pixel_aligned_scene(height: int, width: int, *, sh_degree: int = 4, seed: int = 1234,
                        n_context: int = 2, baseline: float = 0.5, device="cpu") -> Scene:
    """One Gaussian per ERP pixel for each of ``n_context`` context panoramas
    (camera k sits at x = k * baseline, identity rotation): G = n_context * H * W."""
    gen = torch.Generator(device=device).manual_seed(seed)
    dirs = erp_pixel_dirs(height, width, device=device).reshape(-1, 3)
    means, covs, shs, opacs = [], [], [], []
    for k in range(n_context):
        n = dirs.shape[0]
        depth = torch.exp(torch.rand(n, generator=gen, device=device) * (math.log(10.0) - math.log(0.5)) + math.log(0.5))
        origin = torch.tensor([k * baseline, 0.0, 0.0], device=device)
        cov, sh, opac = _attributes(depth, sh_degree, max(height, width), gen)
        means.append(origin + dirs * depth[:, None])
        covs.append(cov)
        shs.append(sh)
        opacs.append(opac)
    return Scene(torch.cat(means), torch.cat(covs), torch.cat(shs), torch.cat(opacs))


def random_cloud_scene(n: int, *, sh_degree: int = 4, seed: int = 1234, ref_width: int = 1024,
                       depth_range=(0.5, 10.0), device="cpu") -> Scene:
    """``n`` Gaussians uniform in direction around the origin, log-uniform depth."""
    gen = torch.Generator(device=device).manual_seed(seed)
    d = torch.randn(n, 3, generator=gen, device=device)
    d = d / d.norm(dim=-1, keepdim=True)
    lo, hi = depth_range
    depth = torch.exp(torch.rand(n, generator=gen, device=device) * (math.log(hi) - math.log(lo)) + math.log(lo))
    cov, sh, opac = _attributes(depth, sh_degree, ref_width, gen)
    return Scene(d * depth[:, None], cov, sh, opac)


def target_pose(seed: int = 0, *, jitter: float = 0.3, max_yaw_deg: float = 10.0, device="cpu") -> Tensor:
    """Camera-to-world [4,4]: translation U(-jitter, jitter)^3 and a small random yaw (about +y)."""
    gen = torch.Generator().manual_seed(seed)
    t = (torch.rand(3, generator=gen) * 2 - 1) * jitter
    yaw = math.radians(max_yaw_deg) * float(torch.rand(1, generator=gen) * 2 - 1)
    c, s = math.cos(yaw), math.sin(yaw)
    m = torch.eye(4)
    m[:3, :3] = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    m[:3, 3] = t
    return m.to(device)


def trajectory(n_frames: int, seed: int = 0, *, radius: float = 0.3, device="cpu") -> Tensor:
    """[n,4,4] smooth camera path (circle in the x-z plane with a slow yaw), cf. the reference's
    interpolated video poses (/root/reference/src/visualization/camera_trajectory/interpolate_trajectory.py:81-91)."""
    out = []
    for f in range(n_frames):
        a = 2 * math.pi * f / max(n_frames, 1)
        yaw = 0.25 * math.sin(a)
        c, s = math.cos(yaw), math.sin(yaw)
        m = torch.eye(4)
        m[:3, :3] = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
        m[:3, 3] = torch.tensor([radius * math.cos(a), 0.05 * math.sin(2 * a), radius * math.sin(a)])
        out.append(m)
    return torch.stack(out).to(device)


def cov3x3_to_cov6(cov: Tensor) -> Tensor:
    """[...,3,3] -> [...,6] in the (xx,xy,xz,yy,yz,zz) order of cuda_splatting.py:115,123."""
    row, col = torch.triu_indices(3, 3)
    return cov[..., row, col]
