"""SURVEY.md sec. 4b rows "single-Gaussian analytic render" and "equal-depth ties" on the CUDA path itself (the oracle has the
same two tests in tests/test_oracle.py): constants of K1 / K6 that no other implementation is needed to check, and the
tie rule of the stable sorts -- Gaussians at EXACTLY the same depth composite in index order, like upstream's stable CUB sort
(recipe of /root/reference/src/scripts/test_splatter.py:38-65 for the single splat)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _identity_pinhole(H, W, dev, **over):
    from splatter360_b200 import camera
    from splatter360_b200 import rasterizer as R
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None]
    cam = camera.pinhole_camera(torch.eye(4)[None], K, torch.tensor([1.0]), torch.tensor([100.0]))
    kw = dict(image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
              viewmatrix=cam.view_matrix[0].to(dev), projmatrix=cam.full_projection[0].to(dev), sh_degree=4,
              campos=torch.zeros(3, device=dev), prefiltered=False, debug=False)
    kw.update(over)
    return R.GaussianRasterizationSettings(**kw)


def test_single_gaussian_analytic_peak_colour_and_footprint():
    from splatter360_b200 import rasterizer as R
    dev = "cuda"
    H = W = 65
    s2 = 0.05 ** 2
    means = torch.tensor([[0.0, 0.0, 4.0]], device=dev)
    cov6 = torch.tensor([[s2, 0, 0, s2, 0, s2]], device=dev)
    sh = torch.zeros(1, 25, 3, device=dev); sh[0, 0] = torch.tensor([1.0, 0.5, -0.2])
    opac = torch.tensor([0.8], device=dev)
    color, st = R.forward_raw(_identity_pinhole(H, W, dev), means, cov6, opac, sh, None)
    img = color.cpu().numpy()
    rgb = 0.28209479 * np.array([1.0, 0.5, -0.2]) + 0.5                      # degree 0 only: C0 * sh0 + 0.5
    np.testing.assert_allclose(img[:, 32, 32], 0.8 * rgb, rtol=2e-5)         # on-axis -> pixel (W-1)/2, alpha = opacity
    var = (W / 2.0 / 4.0) ** 2 * s2 + 0.3                                    # (f/z)^2 sigma^2 + the 0.3 low-pass
    expect = 0.8 * math.exp(-0.5 / var) * rgb[0]
    np.testing.assert_allclose(img[0, 32, 33], expect, rtol=2e-4)
    np.testing.assert_allclose(img[0, 33, 32], expect, rtol=2e-4)
    assert int(st.radii[0]) == math.ceil(3 * math.sqrt(var))
    # opacity above the 0.99 clamp saturates there
    color, _ = R.forward_raw(_identity_pinhole(H, W, dev), means, cov6, torch.tensor([1.0], device=dev), sh, None)
    np.testing.assert_allclose(color.cpu().numpy()[:, 32, 32], 0.99 * rgb, rtol=2e-5)
    # native ERP: the forward direction lands between the four centre pixels of the panorama, symmetric in all of them
    Hs, Ws = 64, 128
    s = _identity_pinhole(Hs, Ws, dev, projection="erp", projmatrix=torch.eye(4, device=dev), viewmatrix=torch.eye(4, device=dev))
    color, _ = R.forward_raw(s, means, cov6, opac, sh, None)
    e = color.cpu().numpy()[0]
    c = e[Hs // 2 - 1:Hs // 2 + 1, Ws // 2 - 1:Ws // 2 + 1]
    assert c.min() > 0 and np.allclose(c, c[0, 0], rtol=1e-4) and e.max() <= c.max() * (1 + 1e-6)


def test_equal_depth_ties_composite_in_index_order():
    from splatter360_b200 import rasterizer as R
    dev = "cuda"
    H = W = 32
    means = torch.tensor([[0.0, 0.0, 3.0]] * 4, device=dev)
    cov6 = torch.tensor([[1e-2, 0, 0, 1e-2, 0, 1e-2]] * 4, device=dev)
    colors = torch.eye(4, 3, device=dev)                                     # red, green, blue, black
    color, st = R.forward_raw(_identity_pinhole(H, W, dev), means, cov6, torch.full((4,), 0.5, device=dev), None, colors)
    pl = st.point_list.cpu().numpy()[:st.num_rendered]
    assert st.num_rendered % 4 == 0 and np.array_equal(pl.reshape(-1, 4), np.tile(np.arange(4), (st.num_rendered // 4, 1)))
    # pixel nearest to the centre (15.5, 15.5): the same alpha a for all four -> weights a, a(1-a), a(1-a)^2 in index order
    px = color.cpu().numpy()[:, 16, 16]
    a = px[0]
    assert 0.3 < a < 0.5
    np.testing.assert_allclose(px, [a, a * (1 - a), a * (1 - a) ** 2], rtol=1e-5)


def test_many_gaussians_at_one_depth_match_the_oracle_list_exactly():
    """3000 Gaussians on the plane z = 3 of an identity camera: every depth key is the same bit pattern, so the whole order
    is decided by the tie rule (index order) in all four sort passes and in the chunked binning."""
    from helpers import rel_l2
    import oracle
    from splatter360_b200 import rasterizer as R
    dev = "cuda"
    H, W, n = 96, 128, 3000
    g = torch.Generator().manual_seed(5)
    means = torch.cat([(torch.rand(n, 2, generator=g) - 0.5) * 5.0, torch.full((n, 1), 3.0)], dim=1)
    cov6 = torch.tensor([[4e-3, 0, 0, 4e-3, 0, 4e-3]]).repeat(n, 1)
    opac = torch.rand(n, generator=g) * 0.6 + 0.2
    colors = torch.rand(n, 3, generator=g)
    s = _identity_pinhole(H, W, dev, tight_bbox=False)
    color, st = R.forward_raw(s, means.to(dev), cov6.to(dev), opac.to(dev), None, colors.to(dev))
    o = oracle.render(means.numpy(), cov6.numpy(), opac.numpy(), colors=colors.numpy(), H=H, W=W,
                      view=s.viewmatrix.cpu().numpy(), proj=s.projmatrix.cpu().numpy(), campos=np.zeros(3, np.float32))
    assert len(set(o["depth"][o["radii"] > 0].tolist())) == 1                 # really one depth
    assert st.num_rendered == o["num_rendered"]
    assert np.array_equal(st.point_list.cpu().numpy().astype(np.uint32)[:st.num_rendered], o["inst_gid"])
    assert rel_l2(color.cpu().numpy(), o["color"]) < 1e-5
