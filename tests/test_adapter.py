"""Row f4: the fused Gaussian adapter (splatter360_b200/adapter.py, csrc/adapter.cu) against the reference's own
GaussianAdapterERP (/root/reference/src/model/encoder/common/gaussian_adapter_erp.py:49-119), whose outputs AND
input-gradients are committed as tests/golden/adapter.npz (tests/golden/make_golden_adapter.py runs the reference module
unmodified; only e3nn is stubbed with oracle/e3nn_wigner.py -- SH rotation parity is therefore 'unpinned', everything
else is pinned by reference code).

CPU: the kernels' per-Gaussian math compiled for the host (tests/host_harness) vs the golden vector; properties of the
restated Wigner matrices.  GPU: the module through the C-ABI vs the golden vector and vs the torch formulation at size."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

from helpers import rel_l2
from test_host_math import harness  # noqa: F401  (fixture: host build of the kernels' math)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rand_rot(g, n):
    q = torch.randn(n, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    x, y, z, w = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                        2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(n, 3, 3)


def test_restated_wigner_matrices_are_a_representation_of_so3():
    from oracle import e3nn_wigner as W
    g = torch.Generator().manual_seed(0)
    R1, R2 = _rand_rot(g, 6), _rand_rot(g, 6)
    for l in range(5):
        D1, D2 = W.wigner_D(l, *W.matrix_to_angles(R1)), W.wigner_D(l, *W.matrix_to_angles(R2))
        D12 = W.wigner_D(l, *W.matrix_to_angles(R1 @ R2))
        eye = torch.eye(2 * l + 1, dtype=torch.float64)
        assert (D1 @ D2 - D12).abs().max() < 1e-12                      # homomorphism
        assert (D1 @ D1.transpose(-1, -2) - eye).abs().max() < 1e-12      # orthogonal
        w = torch.acos(((R1.diagonal(dim1=-2, dim2=-1).sum(-1) - 1) / 2).clamp(-1, 1))
        char = torch.sin((2 * l + 1) * w / 2) / torch.sin(w / 2)
        assert (D1.diagonal(dim1=-2, dim2=-1).sum(-1) - char).abs().max() < 1e-10   # character of the degree-l irrep
    assert (W.wigner_D(1, *W.matrix_to_angles(R1)) - R1).abs().max() < 1e-12        # e3nn: the l = 1 irrep is (x, y, z)


def test_product_sh_rotation_blocks_equal_the_restatement():
    from oracle import e3nn_wigner as W
    from splatter360_b200 import adapter
    R = _rand_rot(torch.Generator().manual_seed(1), 5)
    mask = torch.ones(25)
    for d in range(1, 5):
        mask[d * d:(d + 1) ** 2] = 0.1 * 0.25 ** d
    blk = adapter.sh_rotation_blocks(R.float(), 4, mask)
    M = W.sh_rotation_blocks(R, 4) * mask.double()[None, None, :]
    off = 0
    for l in range(5):
        k = 2 * l + 1
        assert (blk[:, off:off + k * k].reshape(5, k, k).double() - M[:, l * l:(l + 1) ** 2, l * l:(l + 1) ** 2]).abs().max() < 1e-6
        off += k * k


def _golden():
    g = np.load(os.path.join(GOLD, "adapter.npz"))
    h, w = int(g["h"]), int(g["w"])
    b, v = g["depths"].shape[:2]
    return g, h, w, b, v


def test_kernel_math_on_the_host_matches_the_reference_adapter(harness):  # noqa: F811
    from splatter360_b200 import adapter
    g, h, w, b, v = _golden()
    G = b * v * h * w
    ext = torch.from_numpy(g["extrinsics"]).reshape(b * v, 4, 4)
    pose = np.ascontiguousarray(torch.cat([ext[:, :3, :3].reshape(b * v, 9), ext[:, :3, 3]], -1).numpy())
    rot = np.ascontiguousarray(adapter.sh_rotation_blocks(ext[:, :3, :3], 4, torch.from_numpy(g["sh_mask"])).numpy())
    raw = np.ascontiguousarray(g["raw"].reshape(G, -1)); dep = np.ascontiguousarray(g["depths"].reshape(G))
    means = np.zeros((G, 3), np.float32); cov = np.zeros((G, 9), np.float32); sh = np.zeros((G, 75), np.float32)
    gm = np.ascontiguousarray(g["cot_means"].reshape(G, 3)); gc = np.ascontiguousarray(g["cot_cov"].reshape(G, 9))
    gs = np.ascontiguousarray(g["cot_sh"].reshape(G, 75))
    d_raw = np.zeros_like(raw); d_dep = np.zeros_like(dep)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = harness.s360h_adapter(b * v, h, w, 4, ctypes.c_float(float(g["scale_min"])), ctypes.c_float(float(g["scale_max"])), 0,
                               p(raw), p(dep), p(pose), p(rot), p(means), p(cov), p(sh), p(gm), p(gc), p(gs), p(d_raw), p(d_dep))
    assert rc == 0
    assert rel_l2(means, g["means"].reshape(G, 3)) < 2e-6
    assert rel_l2(cov, g["covariances"].reshape(G, 9)) < 5e-6
    assert rel_l2(sh, g["harmonics"].reshape(G, 75)) < 5e-6
    assert rel_l2(d_raw, g["d_raw"].reshape(G, -1)) < 2e-5        # hand-derived backward vs the reference module's autograd
    assert rel_l2(d_dep, g["d_depths"].reshape(G)) < 2e-5         # (the reference's means carry no gradient: no_grad)


def _torch_adapter(ext, depths, raw, h, w, smin, smax, sh_degree, mask):
    """The reference's op sequence (gaussian_adapter_erp.py:61-119) in float64 torch, for sizes the golden file does not hold."""
    from oracle import e3nn_wigner as W
    from splatter360_b200 import camera, synthetic
    d_sh = (sh_degree + 1) ** 2
    scales, rot, sh = raw.split((3, 4, 3 * d_sh), dim=-1)
    scales = (smin + (smax - smin) * scales.sigmoid()) * depths[..., None] / max(h, w)
    rot = rot / (rot.norm(dim=-1, keepdim=True) + 1e-8)
    sh = sh.reshape(*sh.shape[:-1], 3, d_sh) * mask
    Rq = synthetic.quaternion_to_matrix(rot)
    cov = Rq @ torch.diag_embed(scales * scales) @ Rq.transpose(-1, -2)
    Rc = ext[:, None, :3, :3]
    cov = Rc @ cov @ Rc.transpose(-1, -2)
    with torch.no_grad():
        dirs = camera.erp_pixel_dirs(h, w).reshape(-1, 3).to(raw)
    means = (Rc @ (dirs[None] * depths[..., None].detach())[..., None])[..., 0] + ext[:, None, :3, 3]
    harm = W.rotate_sh(sh, ext[:, None, None, :3, :3])
    return means, cov, harm


@pytest.mark.gpu
def test_adapter_module_matches_the_reference_golden_vector_on_gpu():
    from splatter360_b200 import adapter
    g, h, w, b, v = _golden()
    dev = "cuda"
    mod = adapter.GaussianAdapterERP(adapter.GaussianAdapterERPCfg(float(g["scale_min"]), float(g["scale_max"]), 4)).to(dev)
    depths = torch.from_numpy(g["depths"]).to(dev).requires_grad_()
    raw = torch.from_numpy(g["raw"]).to(dev).requires_grad_()
    ext = torch.from_numpy(g["extrinsics"]).to(dev)[:, :, None, None, None]
    out = mod("hm3d", ext, depths, torch.from_numpy(g["opacities"]).to(dev), raw, (h, w))
    for k, tol in (("means", 2e-6), ("covariances", 5e-6), ("harmonics", 5e-6), ("scales", 2e-6), ("rotations", 2e-6)):
        assert getattr(out, k).shape == g[k].shape, k
        assert rel_l2(getattr(out, k).detach().cpu().numpy(), g[k]) < tol, k
    assert torch.equal(out.opacities.cpu(), torch.from_numpy(g["out_opacities"]))
    ((out.means * torch.from_numpy(g["cot_means"]).to(dev)).sum() + (out.covariances * torch.from_numpy(g["cot_cov"]).to(dev)).sum()
     + (out.harmonics * torch.from_numpy(g["cot_sh"]).to(dev)).sum()).backward()
    assert rel_l2(raw.grad.cpu().numpy(), g["d_raw"]) < 2e-5
    assert rel_l2(depths.grad.cpu().numpy(), g["d_depths"]) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("b,v,h,w,deg", [(1, 2, 64, 128, 4), (2, 1, 24, 40, 2), (1, 1, 17, 23, 0)])
def test_adapter_module_matches_the_torch_formulation_at_size(b, v, h, w, deg):
    """Ragged sizes (the last CTA is partial and its block is not 16-byte sized: no TMA bulk copy there), lower SH degrees."""
    from splatter360_b200 import adapter
    dev = "cuda"
    gen = torch.Generator().manual_seed(3)
    d_sh = (deg + 1) ** 2
    r = h * w
    ext = torch.eye(4).repeat(b * v, 1, 1)
    ext[:, :3, :3] = _rand_rot(gen, b * v).float()
    ext[:, :3, 3] = torch.randn(b * v, 3, generator=gen)
    depths = 0.5 + 5 * torch.rand(b * v, r, generator=gen)
    raw = torch.randn(b * v, r, 7 + 3 * d_sh, generator=gen)
    mod = adapter.GaussianAdapterERP(adapter.GaussianAdapterERPCfg(0.5, 15.0, deg)).to(dev)
    dd = depths.double().requires_grad_(); rr = raw.double().requires_grad_()
    m_ref, c_ref, s_ref = _torch_adapter(ext.double(), dd, rr, h, w, 0.5, 15.0, deg, mod.sh_mask.cpu().double())
    cm, cc, cs = (torch.randn(t.shape, generator=gen) for t in (m_ref, c_ref, s_ref))
    ((m_ref * cm).sum() + (c_ref * cc).sum() + (s_ref * cs).sum()).backward()
    d_g = depths.to(dev).reshape(b, v, r, 1, 1).requires_grad_(); r_g = raw.to(dev).reshape(b, v, r, 1, 1, -1).requires_grad_()
    out = mod("hm3d", ext.to(dev).reshape(b, v, 1, 1, 1, 4, 4), d_g, torch.ones(b, v, r, 1, 1, device=dev), r_g, (h, w))
    assert rel_l2(out.means.detach().cpu().reshape(b * v, r, 3).numpy(), m_ref.detach().numpy()) < 2e-6
    assert rel_l2(out.covariances.detach().cpu().reshape(b * v, r, 3, 3).numpy(), c_ref.detach().numpy()) < 5e-6
    assert rel_l2(out.harmonics.detach().cpu().reshape(b * v, r, 3, d_sh).numpy(), s_ref.detach().numpy()) < 5e-6
    ((out.means.reshape(b * v, r, 3) * cm.to(dev)).sum() + (out.covariances.reshape(b * v, r, 3, 3) * cc.to(dev)).sum()
     + (out.harmonics.reshape(b * v, r, 3, d_sh) * cs.to(dev)).sum()).backward()
    assert rel_l2(r_g.grad.cpu().reshape(b * v, r, -1).numpy(), rr.grad.numpy()) < 2e-5
    assert rel_l2(d_g.grad.cpu().reshape(b * v, r).numpy(), dd.grad.numpy()) < 2e-5
