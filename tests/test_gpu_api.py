"""GPU tests through the public API: edge cases, the colours/depth path, scale+rotation path, the decoder-level
mirrors, determinism, and size-independent properties at BASELINE.json's full size."""
import numpy as np
import pytest
import torch

from helpers import make_case, make_settings, rel_l2, run_cuda, run_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north_star: <= 1e-4 relative L2


def test_colors_precomp_path_matches_oracle():
    for mode, H, W in (("pinhole", 64, 80), ("erp", 48, 96)):
        case = make_case(1500, mode, H, W, seed=9)
        case["colors"] = torch.rand(1500, 3, generator=torch.Generator().manual_seed(3))
        dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(4))
        o = run_oracle(case, dL=dL, use_sh=False)
        c = run_cuda(case, dL=dL, use_sh=False)
        assert rel_l2(c["color"], o["color"]) < TOL
        for k in ("d_means", "d_cov6", "d_opac", "d_colors", "d_means2D"):
            assert rel_l2(c[k], o[k]) < TOL, (mode, k)


def test_lower_sh_degrees_and_odd_coefficient_counts():
    for deg, M in ((0, 1), (1, 4), (2, 9), (3, 16), (3, 25)):
        case = make_case(800, "pinhole", 48, 64, seed=deg, sh_degree=4)
        case["shs"] = case["shs"][:, :M].contiguous()
        case["sh_degree"] = deg
        dL = torch.randn(3, 48, 64, generator=torch.Generator().manual_seed(4))
        o = run_oracle(case, dL=dL)
        c = run_cuda(case, dL=dL)
        assert rel_l2(c["color"], o["color"]) < TOL
        assert rel_l2(c["d_shs"], o["d_shs"]) < TOL
        assert c["d_shs"].shape == (800, M, 3)


def test_stock_upstream_degree_cap():
    """max_sh_degree=3 reproduces stock upstream (degree-4 coefficients ignored, zero gradient)."""
    case = make_case(500, "pinhole", 48, 64, seed=2)
    dL = torch.randn(3, 48, 64, generator=torch.Generator().manual_seed(4))
    o = run_oracle(case, dL=dL, max_sh_degree=3)
    c = run_cuda(case, dL=dL, max_sh_degree=3)
    assert rel_l2(c["color"], o["color"]) < TOL and rel_l2(c["d_shs"], o["d_shs"]) < TOL
    assert np.all(c["d_shs"][:, 16:] == 0)


@pytest.mark.parametrize("mode", ["pinhole", "erp"])
def test_edge_cases(mode):
    from splatter360_b200.rasterizer import GaussianRasterizer
    H, W = (37, 53) if mode == "pinhole" else (37, 64)   # non-multiple-of-16 sizes
    case = make_case(400, mode, H, W, seed=4)
    # a few degenerate Gaussians: behind / at the camera, zero opacity, tiny opacity, huge, repeated depth
    case["means"][0] = torch.tensor([0.0, 0.0, -5.0]) + case["campos"]
    case["means"][1] = case["campos"].clone()
    case["opac"][2] = 0.0
    case["opac"][3] = 1e-3
    case["cov6"][4] = torch.tensor([4.0, 0, 0, 4.0, 0, 4.0])
    case["means"][6] = case["means"][5]
    case["opac"][7] = 0.999   # alpha clamp at 0.99
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(8))
    o = run_oracle(case, dL=dL)
    c = run_cuda(case, dL=dL)
    assert np.array_equal(c["radii"], o["radii"])
    assert rel_l2(c["color"], o["color"]) < TOL
    for k in ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"):
        assert np.isfinite(c[k]).all()
        assert rel_l2(c[k], o[k]) < 2 * TOL, k
    # empty scene and fully culled scene
    s = make_settings(case)
    z = lambda *shape: torch.zeros(*shape, device="cuda")
    img, radii = GaussianRasterizer(s)(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), colors_precomp=z(0, 3), cov3D_precomp=z(0, 6))
    assert img.shape == (3, H, W) and radii.numel() == 0
    assert torch.allclose(img, case["bg"].cuda()[:, None, None].expand(3, H, W))
    far_means = (case["campos"] + torch.tensor([0.0, 0.0, 1e-3])).repeat(10, 1).cuda().requires_grad_()
    img, radii = GaussianRasterizer(s)(means3D=far_means, means2D=torch.zeros_like(far_means), opacities=torch.ones(10, 1, device="cuda"),
                                       colors_precomp=torch.ones(10, 3, device="cuda"), cov3D_precomp=torch.ones(10, 6, device="cuda"))
    assert (radii == 0).all()
    img.sum().backward()
    assert torch.all(far_means.grad == 0)


def test_scales_rotations_path_and_mark_visible():
    from splatter360_b200.rasterizer import GaussianRasterizer, build_covariance_6
    case = make_case(600, "pinhole", 48, 64, seed=6)
    s = make_settings(case)
    g = torch.Generator().manual_seed(0)
    scales = (torch.rand(600, 3, generator=g) * 0.05 + 0.01).cuda().requires_grad_()
    rot = torch.randn(600, 4, generator=g)
    rot = (rot / rot.norm(dim=-1, keepdim=True)).cuda().requires_grad_()
    means = case["means"].cuda()
    kw = dict(means3D=means, means2D=torch.zeros_like(means), opacities=case["opac"].cuda()[:, None], shs=case["shs"].cuda())
    img1, _ = GaussianRasterizer(s)(scales=scales, rotations=rot, **kw)
    img2, _ = GaussianRasterizer(s)(cov3D_precomp=build_covariance_6(scales, rot).detach(), **kw)
    assert torch.equal(img1, img2)
    img1.sum().backward()
    assert scales.grad is not None and rot.grad is not None and torch.isfinite(scales.grad).all()
    vis = GaussianRasterizer(s).markVisible(means)
    w2c = case["view"].cuda()
    z = means @ w2c[:3, 2] + w2c[3, 2]
    assert torch.equal(vis, z > 0.2)


def test_forward_is_deterministic_and_backward_reproducible():
    case = make_case(20000, "erp", 128, 256, seed=12)
    dL = torch.randn(3, 128, 256, generator=torch.Generator().manual_seed(1))
    a, b = run_cuda(case, dL=dL), run_cuda(case, dL=dL)
    assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["radii"], b["radii"])
    for k in ("d_means", "d_cov6", "d_opac", "d_shs"):
        assert rel_l2(a[k], b[k]) < 1e-5   # float atomics: order may differ, values agree to round-off


def test_tight_bbox_is_image_exact():
    """Dropping tiles the alpha >= 1/255 ellipse cannot reach must not change a single pixel."""
    for mode, H, W in (("pinhole", 96, 128), ("erp", 64, 128)):
        case = make_case(4000, mode, H, W, seed=21)
        a = run_cuda(case, tight_bbox=True)["color"]
        b = run_cuda(case, tight_bbox=False)["color"]
        assert np.array_equal(a, b)


def test_decoder_level_api_matches_oracle():
    """render_cuda / render_erp / render_depth_* with the reference's tensor conventions (b g 3 3 covariances,
    b g 3 d_sh harmonics, 1/near rescale)."""
    import oracle
    from splatter360_b200 import camera, decoder, synthetic
    sc = synthetic.random_cloud_scene(3000, sh_degree=4, seed=5, ref_width=64, depth_range=(0.5, 4.0))
    pose = synthetic.target_pose(5)
    near, far = torch.tensor([0.5]), torch.tensor([20.0])
    H, W = 64, 128
    dev = "cuda"
    args = [t[None].to(dev) for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]
    img = decoder.render_erp(pose[None].to(dev), near.to(dev), far.to(dev), (H, W), torch.zeros(1, 3, device=dev), *args)
    # oracle on the rescaled scene (cuda_splatting.py:64-71 semantics)
    s = 1 / near[0]
    pose_s = pose.clone(); pose_s[:3, 3] *= s
    cam = camera.erp_camera(pose_s[None])
    o = oracle.render((sc.means * s).numpy(), synthetic.cov3x3_to_cov6(sc.covariances * s * s).numpy(), sc.opacities.numpy(),
                      shs=sc.harmonics.permute(0, 2, 1).contiguous().numpy(), H=H, W=W, view=cam.view_matrix[0].numpy(),
                      proj=cam.full_projection[0].numpy(), campos=cam.campos[0].numpy(), sh_degree=4, mode="erp", stages=False)
    assert rel_l2(img[0].detach().cpu().numpy(), o["color"]) < TOL
    depth = decoder.render_depth_erp(pose[None].to(dev), near.to(dev), far.to(dev), (H, W), args[0], args[1], args[3])
    assert depth.shape == (1, H, W) and torch.isfinite(depth).all() and depth.max() > 0
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].to(dev)
    dec = decoder.DecoderSplattingCUDA((0.0, 0.0, 0.0)).to(dev)
    out = dec(decoder.Gaussians(*args), pose[None, None].to(dev), K[None], near[None].to(dev), far[None].to(dev), (64, 64), depth_mode="depth")
    assert out.color.shape == (1, 1, 3, 64, 64) and out.depth.shape == (1, 1, 64, 64)
    cam_p = camera.pinhole_camera(pose_s[None], K.cpu(), near * s, far * s)
    op = oracle.render((sc.means * s).numpy(), synthetic.cov3x3_to_cov6(sc.covariances * s * s).numpy(), sc.opacities.numpy(),
                       shs=sc.harmonics.permute(0, 2, 1).contiguous().numpy(), H=64, W=64, view=cam_p.view_matrix[0].numpy(),
                       proj=cam_p.full_projection[0].numpy(), campos=cam_p.campos[0].numpy(), tanfovx=float(cam_p.tan_fov_x[0]),
                       tanfovy=float(cam_p.tan_fov_y[0]), sh_degree=4, stages=False)
    assert rel_l2(out.color[0, 0].detach().cpu().numpy(), op["color"]) < TOL


def test_full_size_properties():
    """BASELINE.json configs[2] size (1,048,576 Gaussians, 512x1024 ERP): properties that need no oracle run.
    (a) determinism, (b) background linearity: image(bg) = image(0) + T_final * bg, (c) the backward pass is
    linear in dL/dcolor, (d) permuting the Gaussians permutes the gradients and leaves the image unchanged."""
    from splatter360_b200 import camera, rasterizer, synthetic
    dev = "cuda"
    H, W = 512, 1024
    sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=77, device=dev)
    cam = camera.erp_camera(synthetic.target_pose(7).to(dev)[None])
    means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()

    def settings(bg):
        return rasterizer.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0, viewmatrix=cam.view_matrix[0],
            projmatrix=cam.full_projection[0], sh_degree=4, campos=cam.campos[0], prefiltered=False, debug=False, projection="erp")
    zero = torch.zeros(3, device=dev)
    img0, st0 = rasterizer.forward_raw(settings(zero), means, cov6, op, shs, None)
    img0b, _ = rasterizer.forward_raw(settings(zero), means, cov6, op, shs, None)
    assert torch.equal(img0, img0b)
    assert st0.num_visible > 1_000_000 and st0.num_rendered > st0.num_visible
    bg = torch.tensor([0.3, 0.6, 0.9], device=dev)
    img1, st1 = rasterizer.forward_raw(settings(bg), means, cov6, op, shs, None)
    # final transmittance: recover from a white-minus-black pair
    imgw, _ = rasterizer.forward_raw(settings(torch.ones(3, device=dev)), means, cov6, op, shs, None)
    T = (imgw - img0).mean(0)
    assert torch.allclose(img1, img0 + T[None] * bg[:, None, None], atol=2e-6)
    g = torch.Generator(device=dev).manual_seed(0)
    d1 = torch.randn(3, H, W, device=dev, generator=g) / (3 * H * W)
    d2 = torch.randn(3, H, W, device=dev, generator=g) / (3 * H * W)
    s0 = settings(zero)
    g1 = rasterizer.backward_raw(s0, means, cov6, op, shs, None, st0, d1)
    g2 = rasterizer.backward_raw(s0, means, cov6, op, shs, None, st0, d2)
    g12 = rasterizer.backward_raw(s0, means, cov6, op, shs, None, st0, 2.0 * d1 - 0.5 * d2)
    for k in ("means3D", "cov3D", "opacities", "shs", "means2D"):
        lin = 2.0 * g1[k] - 0.5 * g2[k]
        assert float((g12[k] - lin).norm() / lin.norm()) < 1e-4, k
    perm = torch.randperm(means.shape[0], device=dev, generator=g)
    imgp, stp = rasterizer.forward_raw(s0, means[perm].contiguous(), cov6[perm].contiguous(), op[perm].contiguous(), shs[perm].contiguous(), None)
    assert float((imgp - img0).norm() / img0.norm()) < 1e-5     # equal-depth ties may reorder, nothing else
    gp = rasterizer.backward_raw(s0, means[perm].contiguous(), cov6[perm].contiguous(), op[perm].contiguous(), shs[perm].contiguous(), None, stp, d1)
    assert float((gp["opacities"] - g1["opacities"][perm]).norm() / g1["opacities"].norm()) < 1e-4
    assert float((gp["shs"] - g1["shs"][perm]).norm() / g1["shs"].norm()) < 1e-4


def test_decoder_layouts_and_scale_gradients_match_oracle():
    """The decoder-level path hands the kernels the reference's raw layouts (harmonics [g,3,d_sh], covariances
    [g,3,3]) plus the 1/near scale; gradients must equal autograd through the reference's explicit copies:
    s * dL/dmeans, s^2 * dL/dcov6 on the upper triangle (zero below), transposed dL/dSH."""
    import oracle
    from splatter360_b200 import camera, decoder, synthetic
    for proj in ("erp", "pinhole"):
        sc = synthetic.random_cloud_scene(2500, sh_degree=4, seed=15, ref_width=64, depth_range=(0.5, 4.0))
        pose = synthetic.target_pose(15)
        near, far = torch.tensor([0.4]), torch.tensor([30.0])
        H, W = (64, 128) if proj == "erp" else (64, 64)
        dev = "cuda"
        m = sc.means[None].to(dev).requires_grad_(); c = sc.covariances[None].to(dev).requires_grad_()
        h = sc.harmonics[None].to(dev).requires_grad_(); o = sc.opacities[None].to(dev).requires_grad_()
        dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(2))
        K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].to(dev)
        if proj == "erp":
            img = decoder.render_erp(pose[None].to(dev), near.to(dev), far.to(dev), (H, W), torch.zeros(1, 3, device=dev), m, c, h, o)
        else:
            img = decoder.render_cuda(pose[None].to(dev), K, near.to(dev), far.to(dev), (H, W), torch.zeros(1, 3, device=dev), m, c, h, o)
        (img[0] * dL.to(dev)).sum().backward()
        s = float(1 / near[0])
        pose_s = pose.clone(); pose_s[:3, 3] *= s
        if proj == "erp":
            cam = camera.erp_camera(pose_s[None]); tan = (1.0, 1.0)
        else:
            cam = camera.pinhole_camera(pose_s[None], K.cpu(), near * s, far * s); tan = (float(cam.tan_fov_x[0]), float(cam.tan_fov_y[0]))
        ref = oracle.render((sc.means * s).numpy(), synthetic.cov3x3_to_cov6(sc.covariances * s * s).numpy(), sc.opacities.numpy(),
                            shs=sc.harmonics.permute(0, 2, 1).contiguous().numpy(), H=H, W=W, view=cam.view_matrix[0].numpy(),
                            proj=cam.full_projection[0].numpy(), campos=cam.campos[0].numpy(), tanfovx=tan[0], tanfovy=tan[1],
                            sh_degree=4, mode=proj, dL_dpix=dL.numpy(), stages=False)
        assert rel_l2(img[0].detach().cpu().numpy(), ref["color"]) < TOL
        assert rel_l2(m.grad[0].cpu().numpy(), s * ref["d_means"]) < TOL
        assert rel_l2(o.grad[0].cpu().numpy(), ref["d_opac"]) < TOL
        assert rel_l2(h.grad[0].cpu().numpy(), ref["d_shs"].transpose(0, 2, 1)) < TOL
        gc = c.grad[0].cpu()
        row, col = torch.triu_indices(3, 3)
        assert rel_l2(gc[:, row, col].numpy(), s * s * ref["d_cov6"]) < TOL
        assert float(gc[:, 1, 0].abs().max()) == 0 and float(gc[:, 2, 0].abs().max()) == 0 and float(gc[:, 2, 1].abs().max()) == 0


@pytest.mark.parametrize("mode", ["depth", "disparity", "relative_disparity", "log"])
def test_fused_depth_channel_equals_separate_depth_pass(mode):
    """the depth image accumulated as a fourth channel of the colour pass equals the reference's
    second rasterisation with depth-as-colour (cuda_splatting.py:226-269), pinhole and erp."""
    from splatter360_b200 import decoder, synthetic
    sc = synthetic.random_cloud_scene(3000, sh_degree=4, seed=25, ref_width=64, depth_range=(0.5, 4.0))
    pose = synthetic.target_pose(25)
    dev = "cuda"
    near, far = torch.tensor([0.5], device=dev), torch.tensor([20.0], device=dev)
    args = [t[None].to(dev) for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None]
    bg = torch.zeros(1, 3, device=dev)
    with torch.no_grad():
        img, d_fused = decoder.render_cuda(pose[None].to(dev), K, near, far, (64, 80), bg, *args, fused_depth_mode=mode)
        img_ref = decoder.render_cuda(pose[None].to(dev), K, near, far, (64, 80), bg, *args)
        d_ref = decoder.render_depth_cuda(pose[None].to(dev), K, near, far, (64, 80), args[0], args[1], args[3], mode=mode)
        assert torch.equal(img, img_ref)
        assert float((d_fused - d_ref).norm() / d_ref.norm()) < 1e-5
        img, d_fused = decoder.render_erp(pose[None].to(dev), near, far, (64, 128), bg, *args, fused_depth_mode=mode)
        d_ref = decoder.render_depth_erp(pose[None].to(dev), near, far, (64, 128), args[0], args[1], args[3], mode=mode)
        assert float((d_fused - d_ref).norm() / d_ref.norm()) < 1e-5
        dec = decoder.DecoderSplattingCUDA().to(dev)
        out = dec(decoder.Gaussians(*args), pose[None, None].to(dev), K[None], near[None], far[None], (64, 80), depth_mode=mode)
        assert float((out.depth[0, 0] - d_fused.new_tensor(0) - decoder.render_depth_cuda(pose[None].to(dev), K, near, far, (64, 80), args[0], args[1], args[3], mode=mode)[0]).norm()) < 1e-3 * float(out.depth.norm() + 1)


def test_native_erp_agrees_with_six_faces_plus_cube2equirec():
    """the native erp render must show the same panorama as the reference's pipeline -- six 90-degree
    pinhole faces (reference face poses, convert_cubemaps_mp.py:151-193), change_order, Cube2Equirec
    (model_wrapper_erp.py:135-158, 395-398).  The two differ in pixel density and in the projection the EWA footprint
    is linearised in, so this is a PSNR check on a band-limited scene (isotropic Gaussians of ~3 degrees); a wrong face
    order, flip or axis convention gives 12-15 dB."""
    from splatter360_b200 import cubemap, decoder
    dev = "cuda"
    H, W, Fw = 256, 512, 128
    n = 1500
    g = torch.Generator().manual_seed(3)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    depth = 1.5 + 4 * torch.rand(n, generator=g)
    means = d * depth[:, None]
    cov = torch.diag_embed(((0.05 * depth) ** 2)[:, None].expand(n, 3))
    sh = torch.zeros(n, 3, 1)
    sh[:, :, 0] = (torch.rand(n, 3, generator=g) - 0.5) / 0.28209479177387814
    op = 0.2 + 0.6 * torch.rand(n, generator=g)
    from splatter360_b200 import synthetic
    pose = synthetic.target_pose(41, jitter=0.2, max_yaw_deg=25.0).to(dev)
    args = [t[None].to(dev) for t in (means, cov, sh, op)]
    near, far = torch.tensor([1.0], device=dev), torch.tensor([100.0], device=dev)
    bg = torch.zeros(1, 3, device=dev)
    with torch.no_grad():
        erp = decoder.render_erp(pose[None], near, far, (H, W), bg, *args)[0]
        faces_c2w = cubemap.cube_face_extrinsics(pose)                        # [6,4,4], dataset order [U B L F R D]
        K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None]
        faces = torch.stack([decoder.render_cuda(faces_c2w[k][None], K, near, far, (Fw, Fw), bg, *args)[0] for k in range(6)])
        strip = torch.cat(list(cubemap.change_order(faces)), dim=-1)[None]   # [1,3,f,6f] in [F R B L U D]
        pano = cubemap.Cube2Equirec(Fw, H, W).to(dev)(strip)[0]   # CUDA stitch kernel
    psnr = 10 * np.log10(1.0 / float(((erp - pano) ** 2).mean()))
    assert psnr > 26.0, f"native erp vs cube2equirec PSNR {psnr:.1f} dB"
    band = slice(H // 2 - 32, H // 2 + 32)   # equator: both samplings are close to 1:1 there
    psnr_c = 10 * np.log10(1.0 / float(((erp[:, band] - pano[:, band]) ** 2).mean()))
    assert psnr_c > 32.0, f"centre band PSNR {psnr_c:.1f} dB"


def test_fused_mse_loss_matches_torch():
    from splatter360_b200.loss import mse_loss
    g = torch.Generator().manual_seed(0)
    for shape in ((3, 64, 128), (3, 37, 53), (5,)):
        a = torch.rand(*shape, generator=g).cuda().requires_grad_()
        b = torch.rand(*shape, generator=g).cuda()
        a2 = a.detach().clone().requires_grad_()
        l1 = mse_loss(a, b, 0.7); (l1 * 3.0).backward()
        l2 = 0.7 * ((a2 - b) ** 2).mean(); (l2 * 3.0).backward()
        assert abs(l1.item() - l2.item()) < 1e-6 * max(1.0, abs(l2.item()))
        assert torch.allclose(a.grad, a2.grad, rtol=1e-5, atol=1e-8)


def test_capacity_mode_and_cuda_graph_replay():
    """instance_capacity removes the host read-back; a whole forward+backward is then capturable as one CUDA graph and a
    replay with a new camera / new Gaussians equals the eager result."""
    from splatter360_b200 import camera, rasterizer, synthetic
    from splatter360_b200.graph import GraphedView
    dev = "cuda"
    H, W = 128, 256
    sc = synthetic.random_cloud_scene(10000, sh_degree=4, seed=51, ref_width=256, depth_range=(0.5, 6.0), device=dev)
    means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
    poses = synthetic.trajectory(3, seed=2).to(dev)
    cams = camera.erp_camera(poses)

    def settings(i, cap=None):
        return rasterizer.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
            viewmatrix=cams.view_matrix[i], projmatrix=cams.full_projection[i], sh_degree=4, campos=cams.campos[i],
            prefiltered=False, debug=False, projection="erp", instance_capacity=cap)
    dL = torch.randn(3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    img_e, st_e = rasterizer.forward_raw(settings(1), means, cov6, op, shs, None)
    g_e = rasterizer.backward_raw(settings(1), means, cov6, op, shs, None, st_e, dL)
    cap = 2 * st_e.num_rendered
    # eager capacity mode
    img_c, st_c = rasterizer.forward_raw(settings(1, cap), means, cov6, op, shs, None)
    assert torch.equal(img_c, img_e) and not rasterizer.overflowed(st_c)
    assert rasterizer.instances_needed(st_c) == st_e.num_rendered
    # too small a capacity is reported, never silently wrong
    _, st_small = rasterizer.forward_raw(settings(1, st_e.num_rendered // 2), means, cov6, op, shs, None)
    assert rasterizer.overflowed(st_small)
    # graph: captured on camera 0, replayed on camera 1
    gv = GraphedView(settings(0, cap), means, cov6, op, shs=shs)
    gv.set_camera(cams.view_matrix[1], cams.full_projection[1], cams.campos[1])
    gv.grad_color.copy_(dL)
    color, grads = gv.replay()
    torch.cuda.synchronize()
    assert torch.equal(color, img_e) and not gv.overflowed()
    for k in ("means3D", "cov3D", "opacities", "shs"):
        assert float((grads[k] - g_e[k]).norm() / g_e[k].norm()) < 1e-5, k


@pytest.mark.parametrize("mode", ["depth", "disparity", "relative_disparity", "log"])
def test_fused_depth_channel_is_differentiable_like_the_separate_pass(mode):
    """Gradients through the fused depth channel (render backward carries a fourth channel, the per-Gaussian backward
    differentiates the depth value w.r.t. the mean) equal autograd through the reference's formulation: a second
    rasterisation with depth-as-colour whose colours are torch functions of the means (cuda_splatting.py:226-269).
    Pinhole, erp, and the batched multi-view pass."""
    from splatter360_b200 import decoder, synthetic
    sc = synthetic.random_cloud_scene(2500, sh_degree=4, seed=31, ref_width=64, depth_range=(0.5, 4.0))
    dev = "cuda"
    poses = torch.stack([synthetic.target_pose(31 + k) for k in range(3)]).to(dev)
    near, far = torch.tensor([0.5], device=dev), torch.tensor([20.0], device=dev)
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None]
    bg = torch.zeros(1, 3, device=dev)
    leaves = [t[None].to(dev).requires_grad_() for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]
    gen = torch.Generator().manual_seed(2)

    def grads(loss):
        for t in leaves:
            t.grad = None
        loss.backward()
        return [None if t.grad is None else t.grad.clone() for t in leaves]

    def check(ga, gb, what):
        for a, b, name in zip(ga, gb, ("means", "covariances", "harmonics", "opacities")):
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < TOL, (what, mode, name, rel_l2(a.cpu().numpy(), b.cpu().numpy()))

    # pinhole
    w1 = torch.randn(1, 3, 64, 80, generator=gen).to(dev); w2 = torch.randn(1, 64, 80, generator=gen).to(dev)
    img, dep = decoder.render_cuda(poses[:1], K, near, far, (64, 80), bg, *leaves, fused_depth_mode=mode)
    g_fused = grads((img * w1).sum() + (dep * w2).sum())
    img_r = decoder.render_cuda(poses[:1], K, near, far, (64, 80), bg, *leaves)
    dep_r = decoder.render_depth_cuda(poses[:1], K, near, far, (64, 80), leaves[0], leaves[1], leaves[3], mode=mode)
    check(g_fused, grads((img_r * w1).sum() + (dep_r * w2).sum()), "pinhole")
    # depth alone in the loss (no colour gradient at all)
    img, dep = decoder.render_cuda(poses[:1], K, near, far, (64, 80), bg, *leaves, fused_depth_mode=mode)
    g_d = grads((dep * w2).sum())
    dep_r = decoder.render_depth_cuda(poses[:1], K, near, far, (64, 80), leaves[0], leaves[1], leaves[3], mode=mode)
    g_dr = grads((dep_r * w2).sum())
    for a, b in ((g_d[0], g_dr[0]), (g_d[1], g_dr[1]), (g_d[3], g_dr[3])):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < TOL
    # erp
    w1 = torch.randn(1, 3, 64, 128, generator=gen).to(dev); w2 = torch.randn(1, 64, 128, generator=gen).to(dev)
    img, dep = decoder.render_erp(poses[:1], near, far, (64, 128), bg, *leaves, fused_depth_mode=mode)
    g_fused = grads((img * w1).sum() + (dep * w2).sum())
    img_r = decoder.render_erp(poses[:1], near, far, (64, 128), bg, *leaves)
    dep_r = decoder.render_depth_erp(poses[:1], near, far, (64, 128), leaves[0], leaves[1], leaves[3], mode=mode)
    check(g_fused, grads((img_r * w1).sum() + (dep_r * w2).sum()), "erp")
    # batched: three pinhole views in one pass, depth channel with gradient
    w1 = torch.randn(1, 3, 3, 48, 64, generator=gen).to(dev); w2 = torch.randn(1, 3, 48, 64, generator=gen).to(dev)
    nb, fb = near[None].expand(1, 3), far[None].expand(1, 3)
    img, dep = decoder.render_cuda_views(poses[None], K[None].expand(1, 3, 3, 3), nb, fb, (48, 64), bg, *leaves, fused_depth_mode=mode)
    g_fused = grads((img * w1).sum() + (dep * w2).sum())
    loss = 0
    for k in range(3):
        img_r = decoder.render_cuda(poses[k:k + 1], K, near, far, (48, 64), bg, *leaves)
        dep_r = decoder.render_depth_cuda(poses[k:k + 1], K, near, far, (48, 64), leaves[0], leaves[1], leaves[3], mode=mode)
        loss = loss + (img_r * w1[:, k]).sum() + (dep_r * w2[:, k]).sum()
    check(g_fused, grads(loss), "batched")


def test_host_scene_feeder_ring_hands_over_the_right_data():
    """Double-buffered upload through a ring of reusable device buffers: every step sees exactly its own host data, also
    when the consumer attaches autograd state to the handed-over tensors and the slots are recycled."""
    from splatter360_b200.io import HostSceneFeeder
    dev = torch.device("cuda")
    feeder = HostSceneFeeder(dev, depth=2)
    steps = [dict(a=torch.full((1 << 20,), float(i)).pin_memory(), b=torch.arange(8, dtype=torch.float32).add_(i).pin_memory())
             for i in range(7)]
    t = feeder.submit(steps[0])
    sums = []
    for i in range(7):
        d = feeder.get(t)
        if i + 1 < 7:
            t = feeder.submit(steps[i + 1])
        a = d["a"].requires_grad_()
        (a * d["b"].sum()).sum().backward()          # some work on the consumer stream + autograd state on the slot
        assert a.grad is not None and float(a.grad[0]) == float(steps[i]["b"].sum())
        sums.append(float(a.detach().sum()))
    assert sums == [float(i) * (1 << 20) for i in range(7)]


def test_graphed_step_equals_the_eager_forward_loss_backward():
    """GraphedStep: forward + fused MSE loss + backward captured in one CUDA graph, in-place inputs, camera updates between
    replays -- same loss and gradients as the eager autograd path for every pose of a short trajectory."""
    from splatter360_b200 import camera, synthetic
    from splatter360_b200.graph import GraphedStep
    from splatter360_b200.loss import mse_loss
    from splatter360_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev, H, W, n = "cuda", 128, 256, 30000
    sc = synthetic.random_cloud_scene(n, seed=2, ref_width=256, device=dev)
    means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    opac = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
    cams = camera.erp_camera(synthetic.trajectory(4, seed=3).to(dev))
    target = torch.rand(3, H, W, device=dev)
    mk = lambda i: GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
        viewmatrix=cams.view_matrix[i], projmatrix=cams.full_projection[i], sh_degree=4, campos=cams.campos[i],
        prefiltered=False, debug=False, projection="erp")
    step = GraphedStep(mk(0), means, cov6, opac, shs, target)
    for i in range(4):
        step.set_camera(cams.view_matrix[i], cams.full_projection[i], cams.campos[i])
        loss, g = step.replay()
        m, c, o, s = (t.clone().requires_grad_() for t in (means, cov6, opac[:, None], shs))
        color, _ = GaussianRasterizer(mk(i))(means3D=m, means2D=torch.zeros_like(m), shs=s, colors_precomp=None, opacities=o,
                                             cov3D_precomp=c)
        ref = mse_loss(color, target)
        ref.backward()
        assert not step.overflowed()
        assert torch.allclose(loss, ref, rtol=1e-6)
        for a, b in ((g["means3D"], m.grad), (g["cov3D"], c.grad), (g["opacities"], o.grad), (g["shs"], s.grad)):
            assert float((a - b).norm() / b.norm()) < 1e-5   # same kernels; float atomics add in a different order
