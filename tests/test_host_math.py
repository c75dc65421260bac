"""CPU tests of the kernels' own per-Gaussian math.  splatter360_b200/csrc/persplat.cuh (projection, EWA covariance, conic,
extents, tile rectangle, SH -> RGB, geometry backward, depth value) is host + device code: the CUDA kernels inline it, and
tests/host_harness/harness.cu compiles the SAME functions for the host.  Here they are checked against the C oracle (forward)
and against finite differences of themselves (backward) -- no GPU needed, no compute call into libsplatter360."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from helpers import make_case, oracle_kwargs, rel_l2, run_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
pytestmark = pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "host_harness", "harness.cu")
    out = os.path.join(HERE, "host_harness", "libharness.so")
    deps = [src] + [os.path.join(HERE, "..", "splatter360_b200", "csrc", f) for f in ("persplat.cuh", "common.cuh", "render_cull.cuh", "adapter_math.cuh", "camera_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                        "-o", out, src], check=True, capture_output=True)
    return ctypes.CDLL(out)


def _view(case, P, M, tight_bbox=0, scene_scale=1.0):
    from splatter360_b200 import _lib
    keep = [np.ascontiguousarray(case[k].numpy(), dtype=np.float32) for k in ("view", "proj", "campos", "bg")]
    v = _lib.S360View()
    v.P, v.M, v.sh_degree = P, M, case["sh_degree"]
    v.image_height, v.image_width = case["H"], case["W"]
    v.mode = {"pinhole": 0, "erp": 1}[case["mode"]]
    v.max_sh_degree, v.tight_bbox = 4, tight_bbox
    v.tanfovx, v.tanfovy = case["tanfovx"], case["tanfovy"]
    v.near_cull, v.fov_clamp, v.lowpass, v.pole_eps = 0.2, 1.3, 0.3, 1e-3
    v.scene_scale, v.sh_layout, v.cov_layout = scene_scale, 0, 0
    v.viewmatrix, v.projmatrix, v.campos, v.bg = (k.ctypes.data for k in keep)
    return v, keep


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _project(harness, case, **kw):
    P, M = case["means"].shape[0], case["shs"].shape[1]
    v, keep = _view(case, P, M, **kw)
    ins = [np.ascontiguousarray(case[k].numpy(), dtype=np.float32) for k in ("means", "cov6", "opac", "shs")]
    out = dict(xy=np.zeros((P, 2), np.float32), conic_opacity=np.zeros((P, 4), np.float32), depth=np.zeros(P, np.float32),
               radii=np.zeros(P, np.int32), tiles_touched=np.zeros(P, np.uint32), rgb=np.zeros((P, 3), np.float32),
               clamped=np.zeros((P, 3), np.uint8), half_extent=np.zeros((P, 2), np.float32))
    rc = harness.s360h_project(ctypes.byref(v), *[_p(a) for a in ins], *[_p(out[k]) for k in
                               ("xy", "conic_opacity", "depth", "radii", "tiles_touched", "rgb", "clamped", "half_extent")])
    assert rc == 0
    return out


@pytest.mark.parametrize("mode,H,W,n", [("pinhole", 96, 128, 4000), ("erp", 64, 128, 4000), ("pinhole", 37, 53, 900), ("erp", 256, 512, 3000)])
def test_kernel_projection_math_matches_oracle_on_the_host(harness, mode, H, W, n):
    """K1 as the GPU runs it (project_view + sh_to_rgb, tight box off = upstream's rectangle) against the oracle's K1."""
    case = make_case(n, mode, H, W, seed=3)
    o = run_oracle(case)
    c = _project(harness, case)
    vis = o["radii"] > 0
    assert vis.sum() > n // 20
    assert np.array_equal(c["radii"], o["radii"])
    assert np.array_equal(c["tiles_touched"], o["tiles_touched"])
    for k in ("xy", "depth", "conic_opacity", "rgb"):
        assert rel_l2(c[k][vis], o[k][vis]) < 1e-5, k
    assert np.array_equal(c["clamped"][vis], o["clamped"][vis])


def test_tight_box_only_removes_tiles(harness):
    """tight_bbox=1 intersects upstream's rectangle with the alpha >= 1/255 box: never more tiles, same geometry."""
    case = make_case(3000, "pinhole", 96, 128, seed=4)
    a, b = _project(harness, case, tight_bbox=0), _project(harness, case, tight_bbox=1)
    assert (b["tiles_touched"] <= a["tiles_touched"]).all() and (b["tiles_touched"] < a["tiles_touched"]).any()
    assert np.array_equal(a["radii"], b["radii"]) and np.array_equal(a["xy"], b["xy"])


@pytest.mark.parametrize("mode", ["pinhole", "erp"])
def test_kernel_geometry_backward_is_the_derivative_of_the_kernel_projection(harness, mode):
    """view_backward turns the moment sums of the render pass into dL/d(mean, cov).  With moments chosen such that the
    screen-space gradients are (gu, gv, gA, gB, gC), the result must be the derivative of
    F = gu px + gv py + gA cA + gB cB_true... evaluated with the kernel's own forward projection (central differences in
    float32, so the tolerance is loose; the exact check is the GPU parity suite)."""
    H, W, n = (64, 64, 300) if mode == "pinhole" else (64, 128, 40)
    case = make_case(n, mode, H, W, seed=6, inflate=30.0)
    base = _project(harness, case)
    vis = base["radii"] > 0
    if mode == "pinhole":   # stay inside the field of view: the 1.3 tan(fov) clamp deliberately breaks differentiability
        vis &= (np.abs(base["xy"][:, 0] - W / 2) < 0.45 * W) & (np.abs(base["xy"][:, 1] - H / 2) < 0.45 * H)
    idx = np.nonzero(vis)[0][:10]
    assert len(idx) >= 6
    rng = np.random.default_rng(0)
    P, M = n, case["shs"].shape[1]
    mom = rng.standard_normal((P, 5)).astype(np.float32)          # sum q dx, q dy, q dx^2, q dxdy, q dy^2
    acc = np.zeros((P, 9), np.float32)
    acc[:, 3:8] = mom
    v, keep = _view(case, P, M)
    ins = [np.ascontiguousarray(case[k].numpy(), dtype=np.float32) for k in ("means", "cov6", "opac")]
    d_means, d_m2, d_cov = np.zeros((P, 3), np.float32), np.zeros((P, 2), np.float32), np.zeros((P, 6), np.float32)
    assert harness.s360h_view_backward(ctypes.byref(v), *[_p(a) for a in ins], _p(acc), _p(d_means), _p(d_m2), _p(d_cov)) == 0
    op = case["opac"].numpy()
    con = base["conic_opacity"]
    # what the moments mean: screen-space gradients in pixel units
    gu = -op * (con[:, 0] * mom[:, 0] + con[:, 1] * mom[:, 1])
    gv = -op * (con[:, 2] * mom[:, 1] + con[:, 1] * mom[:, 0])
    gA, gB, gC = -0.5 * op * mom[:, 2], -op * mom[:, 3], -0.5 * op * mom[:, 4]

    def F(c2):
        pr = _project(harness, c2)
        return (gu * pr["xy"][:, 0] + gv * pr["xy"][:, 1] + gA * pr["conic_opacity"][:, 0] + gB * pr["conic_opacity"][:, 1]
                + gC * pr["conic_opacity"][:, 2]).astype(np.float64)

    def fd(key, j, i, h):
        cp, cm = dict(case), dict(case)
        cp[key] = case[key].clone(); cm[key] = case[key].clone()
        cp[key][i, j] += h; cm[key][i, j] -= h
        return (F(cp)[i] - F(cm)[i]) / (2 * h)

    num_m = np.array([[fd("means", j, i, 2e-3) for j in range(3)] for i in idx])
    num_c = np.array([[fd("cov6", j, i, 2e-4 * float(case["cov6"][i].abs().max())) for j in range(6)] for i in idx])
    assert rel_l2(d_means[idx], num_m) < 2e-3
    assert rel_l2(d_cov[idx], num_c) < 3e-2
    assert np.allclose(d_m2[idx, 0], gu[idx] * 0.5 * W, rtol=1e-5) and np.allclose(d_m2[idx, 1], gv[idx] * 0.5 * H, rtol=1e-5)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_depth_value_and_its_derivative(harness, mode):
    near, far, inv_scale = 0.7, 30.0, 1 / 1.6
    z = np.linspace(0.3, 60.0, 400).astype(np.float32)
    val, grad = np.zeros_like(z), np.zeros_like(z)
    assert harness.s360h_depth_value(mode, ctypes.c_float(inv_scale), ctypes.c_float(near), ctypes.c_float(far), len(z), _p(z), _p(val), _p(grad)) == 0
    zz = z.astype(np.float64) * inv_scale
    eps = 1e-10
    ref = [zz, 1 / zz, 1 - (1 / (zz + eps) - 1 / (far + eps)) / (1 / (near + eps) - 1 / (far + eps) + eps),
           np.log(np.maximum(np.minimum(zz, near), far))][mode]
    assert np.allclose(val, ref, rtol=2e-6, atol=1e-7)
    dn, df = 1 / (near + eps), 1 / (far + eps)
    dref = [np.full_like(zz, inv_scale), -inv_scale / zz ** 2, inv_scale / (zz + eps) ** 2 / (dn - df + eps),
            np.where((zz <= near) & (np.minimum(zz, near) >= far), inv_scale / zz, 0.0)][mode]
    assert np.allclose(grad, dref, rtol=2e-6, atol=1e-9)


def test_block_cull_never_drops_a_contributing_instance(harness):
    """The render kernels evaluate an instance only for the 8x8 pixel blocks that rect_can_contribute lets through.  Brute
    force over the 64 pixels (float64): whenever some pixel of the block sees alpha >= 1/255 the test must say yes
    (image exactness); and it must not be lax -- blocks it lets through without any such pixel are only near misses, and
    distant Gaussians are rejected."""
    rng = np.random.default_rng(7)
    n = 200000
    # random screen-space Gaussians: cov2D = R diag(s1^2, s2^2) R^T + 0.3 I, conic = inverse; centres around the block
    s1, s2 = np.exp(rng.uniform(np.log(0.3), np.log(12.0), n)), np.exp(rng.uniform(np.log(0.3), np.log(12.0), n))
    th = rng.uniform(0, np.pi, n)
    c, s_ = np.cos(th), np.sin(th)
    a = c * c * s1 ** 2 + s_ * s_ * s2 ** 2 + 0.3
    b = c * s_ * (s1 ** 2 - s2 ** 2)
    d = s_ * s_ * s1 ** 2 + c * c * s2 ** 2 + 0.3
    det = a * d - b * b
    cA, cB, cC = d / det, -b / det, a / det
    op = rng.uniform(0.004, 1.0, n)
    centre = rng.uniform(-40, 48, (n, 2))
    block_c = np.tile(np.array([[3.5, 3.5]]), (n, 1))           # pixels 0..7 x 0..7
    rec = np.zeros((n, 12), np.float32)
    rec[:, 0:2] = centre; rec[:, 2] = cA; rec[:, 3] = cB; rec[:, 4] = cC; rec[:, 5] = op
    rec[:, 6] = 1.0; rec[:, 7] = 1.0                              # finite half extents: tight culling enabled
    hit = np.zeros(n, np.uint8); ev = np.zeros((n, 4), np.float32)
    assert harness.s360h_cull(n, _p(rec), _p(np.ascontiguousarray(block_c, dtype=np.float32)), _p(hit), _p(ev)) == 0
    px, py = np.meshgrid(np.arange(8.0), np.arange(8.0))
    r32 = rec.astype(np.float64)
    dx = r32[:, 0, None] - px.reshape(1, -1); dy = r32[:, 1, None] - py.reshape(1, -1)
    power = -0.5 * (r32[:, 2, None] * dx * dx + r32[:, 4, None] * dy * dy) - r32[:, 3, None] * dx * dy
    alpha = r32[:, 5, None] * np.exp(power)
    contributes = ((alpha >= 1.0 / 255.0) & (power <= 0)).any(axis=1)
    assert contributes.sum() > n // 20 and (~contributes).sum() > n // 20
    assert hit[contributes].all(), "the block test culled an instance that reaches alpha >= 1/255 inside the block"
    lax = hit.astype(bool) & ~contributes
    # let through without a contributing pixel: only because the best CONTINUOUS point of the block reaches the threshold
    assert lax.sum() < 0.02 * hit.sum()          # measured: 0.17 % of the blocks let through, all with best alpha > 0.85 / 255
    best_alpha_pixel = alpha.max(axis=1)
    assert (best_alpha_pixel[lax] > 0.2 / 255.0).all()


@pytest.mark.parametrize("mode,H,W", [("pinhole", 96, 128), ("erp", 64, 128)])
def test_tight_box_never_excludes_a_visible_pixel(harness, mode, H, W):
    """tight_bbox: K1 drops the tiles outside the axis-aligned box |dx| <= hx, |dy| <= hy.  Outside that box alpha must be
    below 1/255 for certain (float64 brute force on a pixel grid around every visible Gaussian), otherwise the image would
    change."""
    n = 3000
    case = make_case(n, mode, H, W, seed=8)
    pr = _project(harness, case, tight_bbox=1)
    vis = (pr["radii"] > 0) & np.isfinite(pr["half_extent"][:, 0])
    assert vis.sum() > 100
    xy, co, he = pr["xy"][vis].astype(np.float64), pr["conic_opacity"][vis].astype(np.float64), pr["half_extent"][vis].astype(np.float64)
    g = np.arange(-60, 61, dtype=np.float64)
    dx, dy = np.meshgrid(g, g)
    dx, dy = dx.reshape(1, -1), dy.reshape(1, -1)
    power = -0.5 * (co[:, 0, None] * dx * dx + co[:, 2, None] * dy * dy) - co[:, 1, None] * dx * dy
    alpha = co[:, 3, None] * np.exp(power)
    outside = (np.abs(dx) > he[:, 0, None]) | (np.abs(dy) > he[:, 1, None])
    assert (alpha[outside] < 1.0 / 255.0).all()
    # and the box is not lax: just inside its x and y faces some point still reaches 1/255 * 0.9
    inside_edge = (~outside) & ((np.abs(dx) > he[:, 0, None] - 1.5) | (np.abs(dy) > he[:, 1, None] - 1.5))
    small = (he[:, 0] < 55) & (he[:, 1] < 55)
    assert (np.where(inside_edge, alpha, 0).max(axis=1)[small] > 0.5 / 255.0).mean() > 0.9


def test_frustum_reject_never_drops_a_visible_gaussian_on_cube_faces(harness):
    """project_view rejects most (view, Gaussian) pairs with a cheap conservative bound before any covariance math (the
    batched pass projects every Gaussian into six faces).  Large, elongated Gaussians straddling the frustum edges of all six
    cube faces: radii and tile counts must still equal the oracle's, which has no such shortcut."""
    from splatter360_b200 import camera, cubemap, synthetic
    n, F = 6000, 64
    sc = synthetic.random_cloud_scene(n, sh_degree=4, seed=19, ref_width=1024, depth_range=(0.3, 3.0))
    cov = sc.covariances * (40.0 ** 2)
    cov[::3] = cov[::3] * 25.0                                   # a third of them huge
    faces = cubemap.cube_face_extrinsics(synthetic.target_pose(2))
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].repeat(6, 1, 1)
    cam = camera.pinhole_camera(faces, K, torch.ones(6), torch.full((6,), 100.0))
    seen = 0
    for k in range(6):
        case = dict(means=sc.means.contiguous(), cov6=synthetic.cov3x3_to_cov6(cov).contiguous(), opac=sc.opacities.contiguous(),
                    shs=sc.harmonics.permute(0, 2, 1).contiguous(), H=F, W=F, mode="pinhole", sh_degree=4,
                    view=cam.view_matrix[k].contiguous(), proj=cam.full_projection[k].contiguous(), campos=cam.campos[k].contiguous(),
                    tanfovx=float(cam.tan_fov_x[k]), tanfovy=float(cam.tan_fov_y[k]), bg=torch.zeros(3))
        o = run_oracle(case)
        c = _project(harness, case)
        assert np.array_equal(c["radii"], o["radii"]), k
        assert np.array_equal(c["tiles_touched"], o["tiles_touched"]), k
        seen += int((o["radii"] > 0).sum())
    assert seen > n          # most Gaussians are seen by more than one face at this size


def test_invert4x4_on_the_host_matches_the_float64_inverse(harness):
    """camera_math.cuh (the body of s360_invert4x4's kernel): rigid poses, general matrices that need pivoting, a zero
    leading pivot -- within float rounding of numpy's float64 inverse; singular input yields non-finite values, no crash."""
    from splatter360_b200 import synthetic
    rng = np.random.default_rng(3)
    poses = synthetic.trajectory(40, seed=2).numpy().astype(np.float32)
    general = (rng.standard_normal((60, 4, 4)) + 2 * np.eye(4)).astype(np.float32)
    general[7, 0, 0] = 0.0
    perm = np.eye(4, dtype=np.float32)[[2, 0, 3, 1]][None]                    # every pivot needs a row swap
    m = np.ascontiguousarray(np.concatenate([poses, general, perm]))
    out = np.empty_like(m)
    assert harness.s360h_invert4x4(_p(m), _p(out), ctypes.c_int64(m.shape[0])) == 0
    want = np.linalg.inv(m.astype(np.float64))
    err = np.abs(out - want).max(axis=(1, 2)) / np.abs(want).max(axis=(1, 2))
    assert err.max() < 2e-7, err.max()
    assert np.array_equal(out[-1], perm[0].T)
    sing = np.zeros((1, 4, 4), np.float32); o2 = np.empty_like(sing)
    assert harness.s360h_invert4x4(_p(sing), _p(o2), ctypes.c_int64(1)) == 0 and not np.isfinite(o2).all()
