"""CPU tests of the oracle itself: the plain-C restatement against the independent float64 autograd
restatement, analytic cases, and the edge cases listed in SURVEY.md sec. 4b."""
import math

import numpy as np
import pytest
import torch

from helpers import make_case, oracle_kwargs, rel_l2, run_oracle


def _torch_oracle(case, dL, **over):
    from oracle import torch_oracle
    kw = oracle_kwargs(case, **over)
    m = case["means"].double().requires_grad_()
    c = case["cov6"].double().requires_grad_()
    o = case["opac"].double().requires_grad_()
    s = case["shs"].double().requires_grad_()
    col, aux = torch_oracle.render(m, c, o, shs=s, **kw)
    (col * dL.double()).sum().backward()
    return col.detach().numpy(), aux, dict(d_means=m.grad.numpy(), d_cov6=c.grad.numpy(), d_opac=o.grad.numpy(),
                                           d_shs=s.grad.numpy(), d_means2D=aux["means2D"].grad.numpy())


@pytest.mark.parametrize("mode,H,W", [("pinhole", 48, 64), ("erp", 32, 64), ("pinhole", 40, 50)])
def test_c_oracle_matches_autograd_oracle(mode, H, W):
    case = make_case(120, mode, H, W, seed=5)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(2))
    o = run_oracle(case, dL=dL)
    col, aux, g = _torch_oracle(case, dL)
    assert rel_l2(o["color"], col) < 2e-6
    assert np.array_equal(o["n_contrib"].astype(np.int64), aux["n_contrib"].numpy())
    for k, v in g.items():
        assert rel_l2(o[k], v) < 5e-6, k


def test_sh_basis_is_orthonormal_up_to_degree_4():
    """Pins the 25 real-SH constants (incl. the degree-4 band the fork is assumed to evaluate)."""
    from oracle.torch_oracle import sh_basis
    n = 200
    # Gauss-Legendre in cos(theta) x uniform in phi integrates degree <= 8 polynomials exactly
    xs, ws = np.polynomial.legendre.leggauss(16)
    phi = (np.arange(n) + 0.5) * 2 * math.pi / n
    z = np.repeat(xs, n); w = np.repeat(ws, n) * (2 * math.pi / n)
    r = np.sqrt(1 - z * z)
    d = torch.tensor(np.stack([r * np.cos(np.tile(phi, 16)), r * np.sin(np.tile(phi, 16)), z], -1))
    B = sh_basis(4, d).numpy()
    gram = (B * w[:, None]).T @ B
    assert np.allclose(gram, np.eye(25), atol=1e-9)


def test_single_gaussian_analytic_peak():
    """On-axis isotropic Gaussian: peak alpha = opacity, colour = C0*sh0 + 0.5, footprint = Sigma2D + 0.3 I
    (recipe of /root/reference/src/scripts/test_splatter.py:38-65)."""
    import oracle
    H = W = 65
    means = np.array([[0, 0, 4.0]], np.float32)
    s2 = 0.05 ** 2
    cov6 = np.array([[s2, 0, 0, s2, 0, s2]], np.float32)
    sh = np.zeros((1, 25, 3), np.float32); sh[0, 0] = [1.0, 0.5, -0.2]
    view = np.eye(4, dtype=np.float32)
    from splatter360_b200 import camera
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None]
    cam = camera.pinhole_camera(torch.eye(4)[None], K, torch.tensor([1.0]), torch.tensor([100.0]))
    o = oracle.render(means, cov6, np.array([0.8], np.float32), shs=sh, H=H, W=W, view=cam.view_matrix[0].numpy(),
                      proj=cam.full_projection[0].numpy(), campos=np.zeros(3, np.float32), tanfovx=1.0, tanfovy=1.0,
                      sh_degree=4)
    # centre pixel: ndc 0 -> pixel (W-1)/2 = 32
    np.testing.assert_allclose(o["xy"][0], [32.0, 32.0], atol=1e-4)
    rgb = 0.28209479 * sh[0, 0] + 0.5
    np.testing.assert_allclose(o["color"][:, 32, 32], 0.8 * rgb, rtol=1e-5)
    fx = W / 2.0
    var = (fx / 4.0) ** 2 * s2 + 0.3
    expect = 0.8 * math.exp(-0.5 * 1.0 / var) * rgb[0]
    np.testing.assert_allclose(o["color"][0, 32, 33], expect, rtol=1e-4)
    np.testing.assert_allclose(o["color"][0, 33, 32], expect, rtol=1e-4)


def test_edge_cases_cull_and_thresholds():
    import oracle
    H, W = 32, 48
    from splatter360_b200 import camera
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None]
    cam = camera.pinhole_camera(torch.eye(4)[None], K, torch.tensor([1.0]), torch.tensor([100.0]))
    means = np.array([[0, 0, -1.0],      # behind the camera
                      [0, 0, 0.2],       # exactly at the near-cull plane (z <= 0.2 is culled)
                      [0, 0, 0.2001],    # just in front of it
                      [50, 0, 3.0],      # far off screen (rect area 0)
                      [0, 0, 3.0]], np.float32)   # visible, but opacity below 1/255
    cov6 = np.tile(np.array([1e-3, 0, 0, 1e-3, 0, 1e-3], np.float32), (5, 1))
    op = np.array([0.9, 0.9, 0.9, 0.9, 0.003], np.float32)
    col = np.ones((5, 3), np.float32)
    kw = dict(H=H, W=W, view=cam.view_matrix[0].numpy(), proj=cam.full_projection[0].numpy(), campos=np.zeros(3, np.float32),
              bg=(0.2, 0.3, 0.4))
    o = oracle.render(means, cov6, op, colors=col, **kw)
    assert list(o["radii"] > 0) == [False, False, True, False, True]
    o2 = oracle.render(means[[0, 1, 3, 4]], cov6[:4], op[[0, 1, 3, 4]], colors=col[:4], **kw)
    # nothing contributes: image is the background, transmittance 1, no contributors
    assert np.allclose(o2["color"], np.array([0.2, 0.3, 0.4])[:, None, None])
    assert (o2["n_contrib"] == 0).all() and np.allclose(o2["final_T"], 1.0)


def test_equal_depth_ties_keep_index_order():
    import oracle
    from splatter360_b200 import camera
    H = W = 32
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None]
    cam = camera.pinhole_camera(torch.eye(4)[None], K, torch.tensor([1.0]), torch.tensor([100.0]))
    means = np.array([[0, 0, 3.0]] * 4, np.float32)
    cov6 = np.tile(np.array([1e-2, 0, 0, 1e-2, 0, 1e-2], np.float32), (4, 1))
    o = oracle.render(means, cov6, np.full(4, 0.5, np.float32), colors=np.eye(4, 3, dtype=np.float32), H=H, W=W,
                      view=cam.view_matrix[0].numpy(), proj=cam.full_projection[0].numpy(), campos=np.zeros(3, np.float32))
    r = o["tile_ranges"]
    for t in range(r.shape[0]):
        ids = o["inst_gid"][r[t, 0]:r[t, 1]]
        assert list(ids) == sorted(ids)


def test_empty_scene():
    import oracle
    o = oracle.render(np.zeros((0, 3), np.float32), np.zeros((0, 6), np.float32), np.zeros(0, np.float32),
                      colors=np.zeros((0, 3), np.float32), H=20, W=36, view=np.eye(4, dtype=np.float32),
                      proj=np.eye(4, dtype=np.float32), campos=np.zeros(3, np.float32), bg=(1, 0, 0))
    assert o["num_rendered"] == 0 and np.allclose(o["color"][0], 1.0) and np.allclose(o["color"][1:], 0.0)


def test_erp_jacobian_is_derivative_of_reference_projection():
    """Appendix B2: J of the erp mode equals autograd of the reference's point->pixel map."""
    from splatter360_b200 import camera
    t = torch.tensor([[0.7, -0.4, 1.3], [-2.0, 0.5, -0.3]], dtype=torch.float64, requires_grad=True)
    H, W = 512, 1024
    J_auto = torch.stack([torch.autograd.functional.jacobian(lambda p: camera.erp_project(p, H, W), t[i]) for i in range(2)])
    x, y, z = t.detach().unbind(-1)
    su, sv = -W / (2 * math.pi), -H / math.pi
    q = x * x + z * z; rho = q.sqrt(); r2 = q + y * y
    J = torch.stack([torch.stack([su * z / q, torch.zeros_like(x), -su * x / q], -1),
                     torch.stack([-sv * x * y / (rho * r2), sv * rho / r2, -sv * z * y / (rho * r2)], -1)], 1)
    assert torch.allclose(J, J_auto, atol=1e-9)


def test_erp_seam_wrap():
    """A Gaussian straddling theta = +-pi shows up on both image edges with periodic distance."""
    import oracle
    H, W = 32, 64
    means = np.array([[0.001, 0.0, -2.0]], np.float32)   # looks backwards: u ~ -0.5 / W - 0.5
    cov6 = np.array([[0.05, 0, 0, 0.05, 0, 0.05]], np.float32)
    o = oracle.render(means, cov6, np.array([0.9], np.float32), colors=np.ones((1, 3), np.float32), H=H, W=W,
                      view=np.eye(4, dtype=np.float32), proj=np.eye(4, dtype=np.float32), campos=np.zeros(3, np.float32),
                      mode="erp")
    row = o["color"][0, H // 2]
    assert row[0] > 0.3 and row[W - 1] > 0.3 and row[W // 2] == 0.0
    assert abs(row[0] - row[W - 1]) < 0.05


def test_oracle_erp_agrees_with_six_faces_plus_cube2equirec():
    """CPU version of the convention cross-check: the oracle's erp mode vs the oracle's pinhole mode on the reference's six
    face poses, stitched with the (golden-pinned) Cube2Equirec.  Same sign / order / flip conventions or ~13 dB."""
    import oracle
    from splatter360_b200 import camera, cubemap, synthetic
    H, W, Fw = 64, 128, 32
    n = 400
    g = torch.Generator().manual_seed(3)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    depth = 1.5 + 4 * torch.rand(n, generator=g)
    means = (d * depth[:, None]).numpy()
    s2 = (0.08 * depth) ** 2
    cov6 = torch.stack([s2, 0 * s2, 0 * s2, s2, 0 * s2, s2], -1).numpy()
    op = (0.2 + 0.6 * torch.rand(n, generator=g)).numpy()
    col = torch.rand(n, 3, generator=g).numpy()
    pose = synthetic.target_pose(41, jitter=0.2, max_yaw_deg=25.0)

    def rend(c2w, mode, h, w):
        if mode == "erp":
            cam, tan = camera.erp_camera(c2w[None]), (1.0, 1.0)
        else:
            K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None]
            cam = camera.pinhole_camera(c2w[None], K, torch.tensor([1.0]), torch.tensor([100.0]))
            tan = (float(cam.tan_fov_x[0]), float(cam.tan_fov_y[0]))
        o = oracle.render(means, cov6, op, colors=col, H=h, W=w, view=cam.view_matrix[0].numpy(), proj=cam.full_projection[0].numpy(),
                          campos=cam.campos[0].numpy(), tanfovx=tan[0], tanfovy=tan[1], mode=mode, stages=False)
        return torch.from_numpy(o["color"])

    erp = rend(pose, "erp", H, W)
    fc = cubemap.cube_face_extrinsics(pose)
    faces = torch.stack([rend(fc[k], "pinhole", Fw, Fw) for k in range(6)])
    pano = cubemap.Cube2Equirec(Fw, H, W).forward_reference(torch.cat(list(cubemap.change_order(faces)), dim=-1)[None])[0]
    psnr = 10 * math.log10(1.0 / float(((erp - pano) ** 2).mean()))
    assert psnr > 24.0, psnr


@pytest.mark.parametrize("mode", ["pinhole", "erp"])
def test_oracle_size_independent_properties(mode):
    """Properties the GPU tests use at full size, checked on the oracle itself: permutation invariance, background
    linearity  image(bg) = image(0) + T_final * bg,  and linearity of the backward pass in dL/dcolor."""
    H, W = (48, 64) if mode == "pinhole" else (32, 64)
    case = make_case(300, mode, H, W, seed=9)
    base = run_oracle(case, bg=(0.0, 0.0, 0.0))
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(300, generator=g)
    pc = dict(case)
    for k in ("means", "cov6", "opac", "shs"):
        pc[k] = case[k][perm].contiguous()
    p = run_oracle(pc, bg=(0.0, 0.0, 0.0))
    assert rel_l2(p["color"], base["color"]) < 1e-6
    bg = np.array([0.3, 0.6, 0.9], np.float32)
    withbg = run_oracle(case, bg=bg)
    assert np.allclose(withbg["color"], base["color"] + base["final_T"][None] * bg[:, None, None], atol=2e-6)
    d1 = torch.randn(3, H, W, generator=g); d2 = torch.randn(3, H, W, generator=g)
    g1, g2, g12 = (run_oracle(case, dL=d, stages=False) for d in (d1, d2, 2.0 * d1 - 0.5 * d2))
    for k in ("d_means", "d_cov6", "d_opac", "d_shs"):
        assert rel_l2(g12[k], 2.0 * g1[k] - 0.5 * g2[k]) < 2e-5, k


@pytest.mark.parametrize("mode,dmode", [("pinhole", "depth"), ("pinhole", "disparity"), ("pinhole", "relative_disparity"),
                                        ("erp", "depth"), ("erp", "relative_disparity"), ("pinhole", "log")])
def test_c_oracle_depth_channel_matches_autograd_oracle(mode, dmode):
    """Depth channel of the C oracle (value + hand-derived gradient through the blend weights AND through the depth
    value's dependence on the mean) against float64 autograd: the semantics of the reference's depth-as-colour pass
    (cuda_splatting.py:226-269), where the 'colour' is a torch function of the means."""
    from oracle import torch_oracle
    import oracle
    H, W = (40, 48) if mode == "pinhole" else (32, 64)
    case = make_case(100, mode, H, W, seed=9)
    gen = torch.Generator().manual_seed(3)
    dL = torch.randn(3, H, W, generator=gen)
    dD = torch.randn(H, W, generator=gen)
    dk = dict(depth_mode=dmode, depth_near=0.7, depth_far=30.0, depth_scale=1.6)
    kw = oracle_kwargs(case)
    o = oracle.render(case["means"].numpy(), case["cov6"].numpy(), case["opac"].numpy(), shs=case["shs"].numpy(),
                      dL_dpix=dL.numpy(), dL_ddepth=dD.numpy(), stages=False, **kw, **dk)
    m = case["means"].double().requires_grad_(); c = case["cov6"].double().requires_grad_()
    op = case["opac"].double().requires_grad_(); s = case["shs"].double().requires_grad_()
    col, aux = torch_oracle.render(m, c, op, shs=s, **kw, **dk)
    ((col * dL.double()).sum() + (aux["depth"] * dD.double()).sum()).backward()
    assert rel_l2(o["color"], col.detach().numpy()) < 2e-6
    assert rel_l2(o["depth_image"], aux["depth"].detach().numpy()) < 2e-6
    for k, v in (("d_means", m.grad), ("d_cov6", c.grad), ("d_opac", op.grad), ("d_shs", s.grad)):
        assert rel_l2(o[k], v.numpy()) < 5e-6, k
    # depth alone in the loss
    o2 = oracle.render(case["means"].numpy(), case["cov6"].numpy(), case["opac"].numpy(), shs=case["shs"].numpy(),
                       dL_ddepth=dD.numpy(), stages=False, **kw, **dk)
    for t in (m, c, op, s):
        t.grad = None
    col, aux = torch_oracle.render(m, c, op, shs=s, **kw, **dk)
    (aux["depth"] * dD.double()).sum().backward()
    for k, v in (("d_means", m.grad), ("d_cov6", c.grad), ("d_opac", op.grad)):
        assert rel_l2(o2[k], v.numpy()) < 5e-6, k


def test_depth_channel_analytic_constant_depth():
    """All Gaussians on one plane z = z0 in front of a pinhole camera: the blended depth is z0 * (1 - T_final) exactly
    (weights sum to 1 - T_final), disparity likewise with 1/z0."""
    import oracle
    H, W, n, z0 = 32, 48, 60, 2.5
    g = torch.Generator().manual_seed(4)
    means = torch.cat([(torch.rand(n, 2, generator=g) - 0.5) * 3.0, torch.full((n, 1), z0)], -1)
    cov6 = torch.tensor([0.02, 0, 0, 0.02, 0, 0.02]).repeat(n, 1)
    opac = 0.2 + 0.6 * torch.rand(n, generator=g)
    colors = torch.rand(n, 3, generator=g)
    view = np.eye(4, dtype=np.float32)
    proj = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, -0.1, 0]], dtype=np.float32)   # w_clip = z
    for dmode, val in (("depth", z0), ("disparity", 1 / z0)):
        o = oracle.render(means.numpy(), cov6.numpy(), opac.numpy(), colors=colors.numpy(), H=H, W=W, view=view, proj=proj,
                          campos=np.zeros(3, np.float32), depth_mode=dmode, stages=True)
        assert (1 - o["final_T"]).max() > 0.3
        assert np.allclose(o["depth_image"], val * (1 - o["final_T"]), rtol=2e-5, atol=1e-6)


def test_float32_noise_floor_of_the_gradients(tmp_path):
    """How closely can two faithful float32 implementations agree?  The oracle source compiled twice -- as shipped, and with
    FMA contraction (what nvcc does to the CUDA kernels) -- on a pixel-aligned scene: with a WHITE-NOISE seed gradient the
    per-Gaussian moment sums (q dx, q dx^2, ... over a ~3-pixel footprint) cancel almost completely and one ulp in the
    projected centre moves the gradients by ~1e-4 at the scale of config 3; with an image-like seed they agree to ~1e-5.
    This is why tests/test_gpu_baseline_configs.py holds the north_star bound (1e-4) with the image-like seed."""
    import ctypes
    import subprocess
    import oracle
    from splatter360_b200 import camera, synthetic
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU without FMA")
    so = str(tmp_path / "liboracle_fma.so")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=fast", "-fopenmp", "-shared", "-fPIC", "-o", so, oracle._SRC, "-lm"], check=True)
    H, W = 128, 256
    sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237)
    cam = camera.erp_camera(synthetic.trajectory(8, seed=0)[3][None])
    args = (sc.means.numpy(), synthetic.cov3x3_to_cov6(sc.covariances).numpy(), sc.opacities.numpy())
    kw = dict(shs=sc.harmonics.permute(0, 2, 1).contiguous().numpy(), H=H, W=W, view=cam.view_matrix[0].numpy(),
              proj=cam.full_projection[0].numpy(), campos=cam.campos[0].numpy(), sh_degree=4, mode="erp", stages=False)
    g = torch.Generator().manual_seed(0)
    white = (torch.randn(3, H, W, generator=g) / (3 * H * W)).numpy()
    lo = torch.randn(1, 3, H // 16, W // 16, generator=g)
    smooth = (torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)[0] / (3 * H * W)).numpy()
    base = oracle.lib()
    fma = ctypes.CDLL(so)
    fma.oracle_render_ex.restype = ctypes.c_int
    res = {}
    try:
        for name, L in (("base", base), ("fma", fma)):
            oracle._lib = L
            res[name] = (oracle.render(*args, dL_dpix=white, **kw), oracle.render(*args, dL_dpix=smooth, **kw))
    finally:
        oracle._lib = base
    worst_white = max(rel_l2(res["fma"][0][k], res["base"][0][k]) for k in ("d_means", "d_cov6", "d_means2D"))
    worst_smooth = max(rel_l2(res["fma"][1][k], res["base"][1][k]) for k in ("d_means", "d_cov6", "d_means2D", "d_opac", "d_shs"))
    assert rel_l2(res["fma"][0]["color"], res["base"][0]["color"]) < 2e-5
    assert worst_smooth < 3e-5, worst_smooth
    assert worst_white > 2 * worst_smooth, (worst_white, worst_smooth)


@pytest.mark.parametrize("mode,H,W", [("pinhole", 48, 64), ("erp", 32, 64)])
def test_float64_build_of_the_c_oracle_matches_autograd_oracle(mode, H, W):
    """oracle.render(..., f64=True): the same C source with every float as double -- the reference the float32 paths are
    measured against at the BASELINE sizes.  Against float64 autograd it agrees to the float32 rounding of the algorithm's
    constants (0.3f, 1.3f, the SH constants keep their float32 values in the C build)."""
    case = make_case(120, mode, H, W, seed=5)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(2))
    o64 = run_oracle(case, dL=dL, f64=True)
    o32 = run_oracle(case, dL=dL)
    col, aux, g = _torch_oracle(case, dL)
    assert o64["color"].dtype == np.float64
    assert rel_l2(o64["color"], col) < 2e-7
    assert np.array_equal(o64["inst_gid"], o32["inst_gid"]) and np.array_equal(o64["radii"], o32["radii"])
    for k, v in g.items():
        assert rel_l2(o64[k], v) < 2e-6, k
        assert rel_l2(o64[k], v) < rel_l2(o32[k], v), k     # and it is closer to autograd than the float32 build
