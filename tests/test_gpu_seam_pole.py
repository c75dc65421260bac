"""Targeted native-ERP edge cases on the GPU: Gaussians ON the theta = +-pi seam and AT the poles,
not just whatever a random cloud happens to put there.  CUDA vs the C oracle: image, radii, exact instance list, gradients."""
import ctypes
import math

import numpy as np
import pytest
import torch

from helpers import make_settings, rel_l2, run_cuda, run_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _case(means, cov_scale, H, W, seed, opac=None):
    from splatter360_b200 import camera
    n = means.shape[0]
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, 3, 3, generator=g) * cov_scale[:, None, None]
    cov = A @ A.transpose(-1, -2) + (0.05 * cov_scale[:, None, None]) ** 2 * torch.eye(3)
    row, col = torch.triu_indices(3, 3)
    sh = torch.randn(n, 25, 3, generator=g) * 0.2
    sh[:, 0] += 1.0
    cam = camera.erp_camera(torch.eye(4)[None])
    return dict(means=means.contiguous(), cov6=cov[:, row, col].contiguous(),
                opac=(0.15 + 0.8 * torch.rand(n, generator=g)) if opac is None else opac, shs=sh.contiguous(), H=H, W=W,
                mode="erp", sh_degree=4, view=cam.view_matrix[0].contiguous(), proj=cam.full_projection[0].contiguous(),
                campos=cam.campos[0].contiguous(), tanfovx=1.0, tanfovy=1.0, bg=torch.tensor([0.1, 0.2, 0.3]))


def _dirs(theta, phi):
    """unit directions of the reference's sphere camera frame (utils360.py:151-153)"""
    return torch.stack((phi.cos() * theta.sin(), phi.sin(), phi.cos() * theta.cos()), -1)


def _compare(case, H, W, seed, exact_list=True):
    from splatter360_b200 import _lib, rasterizer
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(seed))
    o = run_oracle(case, dL=dL)
    c = run_cuda(case, dL=dL)
    assert np.array_equal(c["radii"], o["radii"])
    assert rel_l2(c["color"], o["color"]) < TOL
    for k in ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"):
        assert rel_l2(c[k], o[k]) < TOL, (k, rel_l2(c[k], o[k]))
    if exact_list:
        dev = "cuda"
        s = make_settings(case, dev, tight_bbox=False)
        _, st = rasterizer.forward_raw(s, case["means"].to(dev), case["cov6"].to(dev), case["opac"].to(dev), case["shs"].to(dev), None)
        assert st.num_rendered == o["num_rendered"]
        assert np.array_equal(st.point_list.cpu().numpy().astype(np.uint32)[: st.num_rendered], o["inst_gid"])
        tiles = ((H + 15) // 16) * ((W + 15) // 16)
        rng = torch.zeros(tiles, 2, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().s360_debug_unpack_image(H, W, ctypes.c_void_p(st.image_state.data_ptr()), None, None,
                                                       ctypes.c_void_p(rng.data_ptr()), None))
        torch.cuda.synchronize()
        assert np.array_equal(rng.cpu().numpy().astype(np.uint32), o["tile_ranges"])
    return o, c


@pytest.mark.parametrize("H,W,n", [(64, 128, 600), (128, 256, 3000)])
def test_gaussians_on_the_seam(H, W, n):
    """Every centre lies within +-2.5 pixel columns of theta = +-pi (u = -0.5 resp. W - 0.5): each splat straddles the
    seam, is binned into the first AND the last tile column and is composited with the periodic pixel distance."""
    g = torch.Generator().manual_seed(1)
    du = (torch.rand(n, generator=g) - 0.5) * 5.0                        # pixel columns away from the seam
    theta = math.pi - du * (2 * math.pi / W)                               # wraps to (-pi, pi] through sin / cos below
    phi = (torch.rand(n, generator=g) - 0.5) * (0.8 * math.pi)
    depth = 0.6 + 3.0 * torch.rand(n, generator=g)
    means = _dirs(theta, phi) * depth[:, None]
    case = _case(means, 0.02 * depth * (128.0 / W), H, W, seed=2)
    o, c = _compare(case, H, W, seed=3)
    # the test is only meaningful if both edge columns are actually covered
    img = o["color"] - np.array([0.1, 0.2, 0.3], np.float32)[:, None, None] * o["final_T"][None]
    assert np.abs(img[:, :, :3]).sum() > 0 and np.abs(img[:, :, -3:]).sum() > 0
    assert (1 - o["final_T"][:, W // 2]).max() == 0.0


def test_wide_gaussians_across_the_seam_need_the_per_pixel_wrap():
    """Splats wider than half the panorama (the render kernels' 'wide' instantiation) centred near the seam."""
    H, W, n = 64, 128, 40
    g = torch.Generator().manual_seed(5)
    theta = math.pi - (torch.rand(n, generator=g) - 0.5) * 0.6
    phi = (torch.rand(n, generator=g) - 0.5) * 1.0
    depth = 0.5 + torch.rand(n, generator=g)
    means = _dirs(theta, phi) * depth[:, None]
    case = _case(means, 0.45 * depth, H, W, seed=6, opac=0.05 + 0.2 * torch.rand(n, generator=g))
    o, _ = _compare(case, H, W, seed=7)
    assert (o["radii"] >= W // 2).sum() >= 5


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_gaussians_at_the_poles(sign):
    """Centres within a few degrees of the pole (|phi| > 85 deg), including some inside the pole_eps clamp
    (rho < 1e-3 r) and one exactly on the axis: horizontal extents blow up (whole tile rows), the Jacobian is evaluated at
    the clamped centre and its gradient masked."""
    H, W, n = 64, 128, 500
    g = torch.Generator().manual_seed(8)
    off = torch.cat([torch.rand(n - 60, generator=g) * math.radians(5.0), torch.rand(59, generator=g) * 5e-4, torch.zeros(1)])
    phi = sign * (math.pi / 2 - off)
    theta = (torch.rand(n, generator=g) * 2 - 1) * math.pi
    depth = 0.6 + 2.0 * torch.rand(n, generator=g)
    means = _dirs(theta, phi) * depth[:, None]
    case = _case(means, 0.015 * depth, H, W, seed=9)
    o, c = _compare(case, H, W, seed=10)
    rows = slice(0, 4) if sign > 0 else slice(H - 4, H)
    assert (1 - o["final_T"][rows]).max() > 0.5          # the polar rows are covered
    assert (o["tiles_touched"] >= W // 16).sum() > 20    # splats spanning a full tile row exist
