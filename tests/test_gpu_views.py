"""Batched multi-view path: V views of the same Gaussians in ONE rasterizer pass must give
what V separate calls give -- per-view images and exact per-view instance lists against the CPU oracle, gradients equal
to the SUM of the oracle's per-view gradients (what autograd produces over the reference's view loop,
/root/reference/src/model/decoder/decoder_splatting_cuda.py:47-59)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north_star: <= 1e-4 relative L2


def _scene(n, seed, inflate=8.0):
    from splatter360_b200 import synthetic
    sc = synthetic.random_cloud_scene(n, sh_degree=4, seed=seed, ref_width=1024, depth_range=(0.5, 4.0))
    return dict(means=sc.means.contiguous(), cov6=synthetic.cov3x3_to_cov6(sc.covariances * inflate ** 2).contiguous(),
                opac=sc.opacities.contiguous(), shs=sc.harmonics.permute(0, 2, 1).contiguous())


def _cube_cameras(seed, near=1.0, far=100.0):
    """Six 90-degree faces around one panorama pose, the reference's face convention (cubemap.cube_face_extrinsics)."""
    from splatter360_b200 import camera, cubemap, synthetic
    pose = synthetic.target_pose(seed)
    faces = cubemap.cube_face_extrinsics(pose)
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].repeat(6, 1, 1)
    return camera.pinhole_camera(faces, K, torch.full((6,), near), torch.full((6,), far))


def _erp_cameras(n_views, seed):
    from splatter360_b200 import camera, synthetic
    poses = torch.stack([synthetic.target_pose(seed + 7 * k) for k in range(n_views)])
    return camera.erp_camera(poses)


def _settings(cam, H, W, mode, dev, **over):
    from splatter360_b200.rasterizer import GaussianRasterizationSettings
    kw = dict(image_height=H, image_width=W, tanfovx=float(cam.tan_fov_x[0]), tanfovy=float(cam.tan_fov_y[0]),
              bg=torch.tensor([0.1, 0.2, 0.3], device=dev), scale_modifier=1.0, viewmatrix=cam.view_matrix.to(dev),
              projmatrix=cam.full_projection.to(dev), sh_degree=4, campos=cam.campos.to(dev), prefiltered=False,
              debug=False, projection=mode)
    kw.update(over)
    return GaussianRasterizationSettings(**kw)


def _oracle_views(sc, cam, H, W, mode, dL=None, use_sh=True, colors=None, stages=False, **over):
    import oracle
    outs = []
    for k in range(cam.view_matrix.shape[0]):
        kw = dict(H=H, W=W, view=cam.view_matrix[k].numpy(), proj=cam.full_projection[k].numpy(),
                  campos=cam.campos[k].numpy(), bg=np.array([0.1, 0.2, 0.3], dtype=np.float32),
                  tanfovx=float(cam.tan_fov_x[k]), tanfovy=float(cam.tan_fov_y[k]), sh_degree=4, mode=mode,
                  dL_dpix=None if dL is None else dL[k].numpy(), stages=stages)
        kw.update(over)
        if use_sh:
            outs.append(oracle.render(sc["means"].numpy(), sc["cov6"].numpy(), sc["opac"].numpy(), shs=sc["shs"].numpy(), **kw))
        else:
            outs.append(oracle.render(sc["means"].numpy(), sc["cov6"].numpy(), sc["opac"].numpy(), colors=colors.numpy(), **kw))
    return outs


def _run_views(sc, settings, dL=None, use_sh=True, colors=None, dev="cuda"):
    from splatter360_b200.rasterizer import rasterize_views
    means = sc["means"].to(dev).requires_grad_()
    cov6 = sc["cov6"].to(dev).requires_grad_()
    opac = sc["opac"].to(dev)[:, None].clone().requires_grad_()
    feat = (sc["shs"] if use_sh else colors).to(dev).requires_grad_()
    color = rasterize_views(means, opac, cov6, settings, shs=feat if use_sh else None, colors_precomp=None if use_sh else feat)
    out = dict(color=color.detach().cpu().numpy())
    if dL is not None:
        (color * dL.to(dev)).sum().backward()
        out.update(d_means=means.grad.cpu().numpy(), d_cov6=cov6.grad.cpu().numpy(),
                   d_opac=opac.grad.reshape(-1).cpu().numpy(), d_feat=feat.grad.cpu().numpy())
    return out


def _check_sum(c, outs, feat_key):
    for k, ok in (("d_means", "d_means"), ("d_cov6", "d_cov6"), ("d_opac", "d_opac"), ("d_feat", feat_key)):
        ref = sum(np.asarray(o[ok], dtype=np.float64) for o in outs)
        assert rel_l2(c[k], ref) < TOL, (k, rel_l2(c[k], ref))


@pytest.mark.parametrize("n,F", [(4000, 64), (20000, 128), (777, 40)])
def test_six_cube_faces_in_one_pass_match_per_face_oracle(n, F):
    sc = _scene(n, seed=21, inflate=1024.0 / (4 * F))
    cam = _cube_cameras(5)
    dL = torch.randn(6, 3, F, F, generator=torch.Generator().manual_seed(4))
    outs = _oracle_views(sc, cam, F, F, "pinhole", dL=dL)
    c = _run_views(sc, _settings(cam, F, F, "pinhole", "cuda"), dL=dL)
    for k in range(6):
        assert rel_l2(c["color"][k], outs[k]["color"]) < TOL, k
    _check_sum(c, outs, "d_shs")


def test_erp_views_with_distinct_camera_centres_match_oracle():
    """Three native-ERP views from different positions: every Gaussian is in every view (pairs = V * P) and the SH
    colour / gradient must be evaluated per view."""
    n, H, W = 3000, 64, 128
    sc = _scene(n, seed=8)
    cam = _erp_cameras(3, seed=2)
    dL = torch.randn(3, 3, H, W, generator=torch.Generator().manual_seed(6))
    outs = _oracle_views(sc, cam, H, W, "erp", dL=dL)
    c = _run_views(sc, _settings(cam, H, W, "erp", "cuda"), dL=dL)
    for k in range(3):
        assert rel_l2(c["color"][k], outs[k]["color"]) < TOL, k
    _check_sum(c, outs, "d_shs")


def test_pinhole_views_with_distinct_camera_centres_and_precomputed_colours():
    from splatter360_b200 import camera, synthetic
    n, H, W = 2500, 48, 80
    sc = _scene(n, seed=12)
    poses = torch.stack([synthetic.target_pose(3 + k, jitter=0.3, max_yaw_deg=40.0) for k in range(4)])
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].repeat(4, 1, 1)
    cam = camera.pinhole_camera(poses, K, torch.ones(4), torch.full((4,), 100.0))
    dL = torch.randn(4, 3, H, W, generator=torch.Generator().manual_seed(9))
    # SH path, distinct centres
    outs = _oracle_views(sc, cam, H, W, "pinhole", dL=dL)
    c = _run_views(sc, _settings(cam, H, W, "pinhole", "cuda"), dL=dL)
    for k in range(4):
        assert rel_l2(c["color"][k], outs[k]["color"]) < TOL, k
    _check_sum(c, outs, "d_shs")
    # colours path
    colors = torch.rand(n, 3, generator=torch.Generator().manual_seed(3))
    outs = _oracle_views(sc, cam, H, W, "pinhole", dL=dL, use_sh=False, colors=colors)
    c = _run_views(sc, _settings(cam, H, W, "pinhole", "cuda"), dL=dL, use_sh=False, colors=colors)
    for k in range(4):
        assert rel_l2(c["color"][k], outs[k]["color"]) < TOL, k
    _check_sum(c, outs, "d_colors")


def test_batched_stage_parity_exact_instance_lists():
    """The pair buffers, the sorted instance list and the tile ranges of the stacked image are EXACTLY the per-face
    lists of the oracle (tight_bbox off so that the lists are upstream's)."""
    from splatter360_b200 import _lib, rasterizer
    lib = _lib.load()
    n, F, dev = 6000, 64, "cuda"
    sc = _scene(n, seed=33, inflate=4.0)
    cam = _cube_cameras(9)
    outs = _oracle_views(sc, cam, F, F, "pinhole", stages=True, tight_bbox=False)
    s = _settings(cam, F, F, "pinhole", dev, tight_bbox=False)
    color, st = rasterizer.forward_views_raw(s, sc["means"].to(dev), sc["cov6"].to(dev), sc["opac"].to(dev),
                                             sc["shs"].to(dev), None, want_radii=True)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    base = torch.zeros(n, dtype=torch.int32, device=dev); mask = torch.zeros(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.s360_debug_unpack_pairs(n, st.pair_capacity, p(st.geom), p(base), p(mask), p(cnt), None))
    gx = gy = F // 16
    rng = torch.zeros(6 * gx * gy, 2, dtype=torch.int32, device=dev)
    _lib.check(lib.s360_debug_unpack_image(6 * F, F, p(st.image_state), None, None, p(rng), None))
    torch.cuda.synchronize()
    base, mask = base.cpu().numpy().astype(np.int64), mask.cpu().numpy().astype(np.uint32)
    # pair slot -> (view, gaussian)
    npairs = int(cnt.item())
    assert npairs == st.num_pairs == sum(int((o["tiles_touched"] > 0).sum()) for o in outs)
    slot_view = np.full(npairs, -1); slot_gid = np.full(npairs, -1)
    for k in range(6):
        has = (mask >> k) & 1 == 1
        assert np.array_equal(has, outs[k]["tiles_touched"] > 0), k
        rank = np.array([bin(int(m) & ((1 << k) - 1)).count("1") for m in mask[has]])
        slot_view[base[has] + rank] = k
        slot_gid[base[has] + rank] = np.nonzero(has)[0]
    assert (slot_view >= 0).all()
    assert st.num_rendered == sum(o["num_rendered"] for o in outs)
    assert np.array_equal(st.radii.cpu().numpy(), np.stack([o["radii"] for o in outs]))
    pl = st.point_list.cpu().numpy().astype(np.int64)[:st.num_rendered]
    rng = rng.cpu().numpy().astype(np.int64).reshape(6, gx * gy, 2)
    off = 0
    for k in range(6):
        nk = outs[k]["num_rendered"]
        seg = pl[off:off + nk]
        assert (slot_view[seg] == k).all(), k
        assert np.array_equal(slot_gid[seg], outs[k]["inst_gid"]), f"face {k}: sorted instance list differs"
        want = outs[k]["tile_ranges"].astype(np.int64)
        want = np.where((want[:, 1] > want[:, 0])[:, None], want + off, 0)
        assert np.array_equal(rng[k], want), k
        off += nk
        assert rel_l2(color[k].cpu().numpy(), outs[k]["color"]) < 1e-5


def test_batched_equals_separate_calls_bitwise_and_decoder_switch():
    """Same kernels, same arithmetic: the batched images are bit-identical to per-view calls; the decoder's
    batched_views switch changes nothing but the number of passes."""
    from splatter360_b200 import cubemap, decoder, synthetic
    dev = "cuda"
    sc = synthetic.random_cloud_scene(5000, sh_degree=4, seed=5, ref_width=256, depth_range=(0.5, 4.0))
    pose = synthetic.target_pose(5)
    faces = cubemap.cube_face_extrinsics(pose)[None].to(dev)            # [1,6,4,4]
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None, None].repeat(1, 6, 1, 1)
    near, far = torch.full((1, 6), 0.5, device=dev), torch.full((1, 6), 20.0, device=dev)
    g = decoder.Gaussians(*[t[None].to(dev).requires_grad_() for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)])
    grads = []
    outs = []
    for batched in (True, False):
        dec = decoder.DecoderSplattingCUDA((0.0, 0.1, 0.2), batched_views=batched).to(dev)
        out = dec(g, faces, K, near, far, (64, 64))
        w = torch.linspace(0, 1, out.color.numel(), device=dev).reshape(out.color.shape)
        for t in (g.means, g.covariances, g.harmonics, g.opacities):
            t.grad = None
        (out.color * w).sum().backward()
        outs.append(out.color.detach().clone())
        grads.append([t.grad.clone() for t in (g.means, g.covariances, g.harmonics, g.opacities)])
    assert torch.equal(outs[0], outs[1])
    for a, b in zip(*grads):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    # evaluation path: fused depth for all six faces in the same pass
    with torch.no_grad():
        o1 = decoder.DecoderSplattingCUDA(batched_views=True).to(dev)(g, faces, K, near, far, (64, 64), depth_mode="depth")
        o2 = decoder.DecoderSplattingCUDA(batched_views=False).to(dev)(g, faces, K, near, far, (64, 64), depth_mode="depth")
    assert torch.equal(o1.color, o2.color) and torch.equal(o1.depth, o2.depth)


def test_batched_edge_cases_and_pair_capacity():
    from splatter360_b200 import rasterizer
    dev = "cuda"
    cam = _cube_cameras(2)
    # one view through the batched path, odd sizes, P not a multiple of the CTA size
    sc = _scene(333, seed=4)
    cam1 = type(cam)(*[t[3:4] for t in cam])
    o = _oracle_views(sc, cam1, 37, 53, "pinhole")
    c = _run_views(sc, _settings(cam1, 37, 53, "pinhole", dev))
    assert rel_l2(c["color"][0], o[0]["color"]) < TOL
    # nothing visible at all: background only
    far_away = dict(sc, means=sc["means"] * 0 + torch.tensor([0.0, 0.0, 0.05]) + cam.campos[0])
    c = _run_views(far_away, _settings(cam, 32, 32, "pinhole", dev))
    assert np.allclose(c["color"], np.array([0.1, 0.2, 0.3], dtype=np.float32)[None, :, None, None])
    # too small a pair capacity is reported, not silently truncated
    sc = _scene(2000, seed=6)
    with pytest.raises(RuntimeError, match="pair_capacity"):
        _run_views(sc, _settings(cam, 64, 64, "pinhole", dev, pair_capacity=100))
    # a sufficient explicit capacity gives the same image as the default
    full = _run_views(sc, _settings(cam, 64, 64, "pinhole", dev))
    tight = _run_views(sc, _settings(cam, 64, 64, "pinhole", dev, pair_capacity=4000))
    assert np.array_equal(full["color"], tight["color"])
    # sync-free capacity mode: overflow flags on the device
    s = _settings(cam, 64, 64, "pinhole", dev, pair_capacity=100, instance_capacity=100000)
    _, st = rasterizer.forward_views_raw(s, sc["means"].to(dev), sc["cov6"].to(dev), sc["opac"].to(dev), sc["shs"].to(dev), None)
    assert int(st.counters[1].item()) & 2


def test_full_size_cube_faces_one_pass_properties():
    """1,048,576 pixel-aligned Gaussians, six 256x256 faces (the reference's panorama at 512x1024): the one-pass images
    equal six separate calls bit for bit, and the gradients agree to summation-order rounding."""
    from splatter360_b200 import camera, cubemap, rasterizer, synthetic
    dev = "cuda"
    sc = synthetic.pixel_aligned_scene(512, 1024, seed=1237, device=dev)
    means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
    pose = synthetic.target_pose(3).to(dev)
    faces = cubemap.cube_face_extrinsics(pose)
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None].repeat(6, 1, 1)
    cam = camera.pinhole_camera(faces, K, torch.ones(6, device=dev), torch.full((6,), 100.0, device=dev))
    F = 256
    s = _settings(cam, F, F, "pinhole", dev)
    dL = torch.randn(6, 3, F, F, device=dev) / (3 * F * F)
    color, st = rasterizer.forward_views_raw(s, means, cov6, op, shs, None)
    g = rasterizer.backward_views_raw(s, means, cov6, op, shs, None, st, dL)
    acc = None
    for k in range(6):
        sk = s._replace(viewmatrix=cam.view_matrix[k], projmatrix=cam.full_projection[k], campos=cam.campos[k])
        ck, stk = rasterizer.forward_raw(sk, means, cov6, op, shs, None)
        assert torch.equal(ck, color[k]), k
        gk = rasterizer.backward_raw(sk, means, cov6, op, shs, None, stk, dL[k])
        acc = {n: gk[n].double() if acc is None else acc[n] + gk[n].double() for n in ("means3D", "cov3D", "opacities", "shs")}
    for n in ("means3D", "cov3D", "opacities", "shs"):
        assert rel_l2(g[n].cpu().numpy(), acc[n].cpu().numpy()) < 1e-5, n


def test_cube2equirec_kernel_matches_golden_and_reference_formulation():
    """The stitch kernel against (a) the committed golden vector generated by the reference's own Cube2Equirec and
    (b) the torch formulation (forward_reference) at full size, forward and backward; and the face-layout entry
    (change_order + strip + stitch in one kernel) against the explicit torch chain."""
    import os
    from splatter360_b200 import cubemap
    dev = "cuda"
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cube2equirec.npz"))
    c2e = cubemap.Cube2Equirec(8, 16, 32).to(dev)
    out = c2e(torch.from_numpy(g["cube"]).to(dev))
    np.testing.assert_allclose(out.cpu().numpy(), g["erp"], atol=1e-6)
    Fw, H, W = 256, 512, 1024
    c2e = cubemap.Cube2Equirec(Fw, H, W).to(dev)
    gen = torch.Generator().manual_seed(0)
    strip = torch.rand(2, 3, Fw, 6 * Fw, generator=gen).to(dev).requires_grad_()
    w = torch.randn(2, 3, H, W, generator=gen).to(dev)
    ref = c2e.forward_reference(strip)
    (ref * w).sum().backward()
    g_ref = strip.grad.clone(); strip.grad = None
    out = c2e(strip)
    (out * w).sum().backward()
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    assert rel_l2(strip.grad.cpu().numpy(), g_ref.cpu().numpy()) < 1e-5
    faces = torch.rand(2, 6, 3, Fw, Fw, generator=gen).to(dev).requires_grad_()
    chain = torch.stack([c2e.forward_reference(torch.cat(list(cubemap.change_order(faces[b])), dim=-1)[None])[0] for b in range(2)])
    (chain * w).sum().backward()
    g_chain = faces.grad.clone(); faces.grad = None
    fused = c2e.from_faces(faces)
    (fused * w).sum().backward()
    assert rel_l2(fused.detach().cpu().numpy(), chain.detach().cpu().numpy()) < 1e-6
    assert rel_l2(faces.grad.cpu().numpy(), g_chain.cpu().numpy()) < 1e-5


def test_erp_decoder_batched_switch_and_views_entry():
    """DecoderSplattingERP: batched_views on/off give the same panoramas (bitwise) and gradients; render_erp_views equals
    per-view render_erp."""
    from splatter360_b200 import decoder, synthetic
    dev = "cuda"
    sc = synthetic.random_cloud_scene(4000, sh_degree=4, seed=15, ref_width=256, depth_range=(0.5, 4.0))
    poses = torch.stack([synthetic.target_pose(20 + k) for k in range(3)])[None].to(dev)     # [1,3,4,4]
    near, far = torch.full((1, 3), 0.5, device=dev), torch.full((1, 3), 20.0, device=dev)
    g = decoder.Gaussians(*[t[None].to(dev).requires_grad_() for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)])
    outs, grads = [], []
    for batched in (True, False):
        dec = decoder.DecoderSplattingERP((0.1, 0.0, 0.2), batched_views=batched).to(dev)
        out = dec(g, poses, None, near, far, (64, 128))
        for t in (g.means, g.covariances, g.harmonics, g.opacities):
            t.grad = None
        (out.color * torch.linspace(0, 1, out.color.numel(), device=dev).reshape(out.color.shape)).sum().backward()
        outs.append(out.color.detach().clone())
        grads.append([t.grad.clone() for t in (g.means, g.covariances, g.harmonics, g.opacities)])
    assert torch.equal(outs[0], outs[1])
    for a, b in zip(*grads):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    with torch.no_grad():
        col, dep = decoder.render_erp_views(poses, near, far, (64, 128), torch.zeros(1, 3, device=dev), g.means, g.covariances,
                                            g.harmonics, g.opacities, fused_depth_mode="disparity")
        for k in range(3):
            c1, d1 = decoder.render_erp(poses[:, k], near[:, k], far[:, k], (64, 128), torch.zeros(1, 3, device=dev), g.means,
                                        g.covariances, g.harmonics, g.opacities, fused_depth_mode="disparity")
            assert torch.equal(col[:, k], c1) and torch.equal(dep[:, k], d1)


def test_many_overlapping_views_trigger_the_exact_pair_capacity_retry():
    """Eight pinhole views that all look at the same cloud: pairs ~ 8 P exceeds the default 2 P + 4096 pair slots, so the
    forward pass re-runs K1 with the exact count it read back -- results must still equal the per-view oracle."""
    from splatter360_b200 import camera, rasterizer, synthetic
    n, H, W, V = 3000, 48, 64, 8
    sc = _scene(n, seed=17)
    # put the whole cloud in front of the cameras
    sc["means"] = sc["means"].clone()
    sc["means"][:, 2] = sc["means"][:, 2].abs() + 1.0
    poses = torch.eye(4)[None].repeat(V, 1, 1)
    poses[:, 0, 3] = torch.linspace(-0.2, 0.2, V)
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None].repeat(V, 1, 1)
    cam = camera.pinhole_camera(poses, K, torch.ones(V), torch.full((V,), 100.0))
    s = _settings(cam, H, W, "pinhole", "cuda")
    dev = "cuda"
    _, st = rasterizer.forward_views_raw(s, sc["means"].to(dev), sc["cov6"].to(dev), sc["opac"].to(dev), sc["shs"].to(dev), None)
    assert st.num_pairs > 2 * n + 4096 and st.pair_capacity == st.num_pairs     # the retry sized the buffers exactly
    dL = torch.randn(V, 3, H, W, generator=torch.Generator().manual_seed(8))
    outs = _oracle_views(sc, cam, H, W, "pinhole", dL=dL)
    c = _run_views(sc, s, dL=dL)
    for k in range(V):
        assert rel_l2(c["color"][k], outs[k]["color"]) < TOL, k
    _check_sum(c, outs, "d_shs")


def test_stitch_kernel_depth_to_distance_matches_reference_golden():
    """from_faces(depth_to_distance=...) = the reference's depth-panorama chain (golden vector generated by the reference's
    own change_order_batch / depth_to_distance_map_batch / Cube2Equirec, tests/golden/make_golden_depth.py); gradient
    against the torch chain."""
    import os
    from splatter360_b200 import cubemap
    dev = "cuda"
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "depth_panorama.npz"))
    f, H, W = int(g["face_w"]), int(g["H"]), int(g["W"])
    k = [float(x) for x in g["fxfycxcy"]]
    c2e = cubemap.Cube2Equirec(f, H, W).to(dev)
    faces = torch.from_numpy(g["faces"]).to(dev)[:, :, None].contiguous().requires_grad_()      # [v,6,1,f,f]
    pano = c2e.from_faces(faces, depth_to_distance=k)
    np.testing.assert_allclose(pano[:, 0].detach().cpu().numpy(), g["pano"], rtol=2e-5, atol=2e-5)
    w = torch.randn(pano.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    (pano * w).sum().backward()
    g_kernel = faces.grad.clone(); faces.grad = None
    fac = cubemap.depth_to_distance_factor(f, *k).to(dev)
    chain = torch.stack([c2e.forward_reference(torch.cat(list(cubemap.change_order(faces[v]) * fac), dim=-1)[None])[0]
                         for v in range(faces.shape[0])])
    (chain * w).sum().backward()
    assert rel_l2(pano.detach().cpu().numpy(), chain.detach().cpu().numpy()) < 1e-6
    assert rel_l2(g_kernel.cpu().numpy(), faces.grad.cpu().numpy()) < 1e-5
