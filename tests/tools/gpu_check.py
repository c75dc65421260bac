"""First-light GPU check: per-stage parity vs the oracle + a coarse timing at bench scale."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_case, rel_l2, run_cuda, run_oracle, make_settings
from splatter360_b200 import rasterizer, synthetic, camera, _lib
import ctypes

def stage_check(mode, H, W, n, seed=7):
    case = make_case(n, mode, H, W, seed=seed)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1))
    o = run_oracle(case, dL=dL)
    # exact-instance-list mode
    dev = "cuda"
    s = make_settings(case, dev, tight_bbox=False)
    means, cov6, op, shs = (case[k].to(dev) for k in ("means", "cov6", "opac", "shs"))
    color, st = rasterizer.forward_raw(s, means, cov6, op, shs, None)
    torch.cuda.synchronize()
    lib = _lib.load()
    P = n
    xy = torch.zeros(P, 2, device=dev); depth = torch.zeros(P, device=dev); conop = torch.zeros(P, 4, device=dev)
    rgb = torch.zeros(P, 3, device=dev); tiles = torch.zeros(P, dtype=torch.int32, device=dev); cl = torch.zeros(P, 3, dtype=torch.uint8, device=dev)
    lib.s360_debug_unpack_geom(P, ctypes.c_void_p(st.geom.data_ptr()), *[ctypes.c_void_p(t.data_ptr()) for t in (xy, depth, conop, rgb, tiles, cl)], None)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    fT = torch.zeros(H, W, device=dev); nc = torch.zeros(H, W, dtype=torch.int32, device=dev); rng = torch.zeros(gx * gy, 2, dtype=torch.int32, device=dev)
    lib.s360_debug_unpack_image(H, W, ctypes.c_void_p(st.image_state.data_ptr()), *[ctypes.c_void_p(t.data_ptr()) for t in (fT, nc, rng)], None)
    torch.cuda.synchronize()
    vis = o["radii"] > 0
    print(f"[{mode} {H}x{W} P={n}] N oracle={o['num_rendered']} cuda={st.num_rendered} vis={vis.sum()} nvis_cuda={st.num_visible}")
    print("  radii eq", np.array_equal(o["radii"], st.radii.cpu().numpy()),
          "tiles eq", np.array_equal(o["tiles_touched"], tiles.cpu().numpy().astype(np.uint32)))
    print("  xy", rel_l2(xy.cpu().numpy()[vis], o["xy"][vis]), "depth", rel_l2(depth.cpu().numpy()[vis], o["depth"][vis]),
          "conic", rel_l2(conop.cpu().numpy()[vis], o["conic_opacity"][vis]), "rgb", rel_l2(rgb.cpu().numpy()[vis], o["rgb"][vis]),
          "clamped eq", np.array_equal(cl.cpu().numpy()[vis], o["clamped"][vis]))
    pl = st.point_list.cpu().numpy().astype(np.uint32)[:st.num_rendered]
    print("  point_list eq", np.array_equal(pl, o["inst_gid"]), "ranges eq", np.array_equal(rng.cpu().numpy().astype(np.uint32), o["tile_ranges"]))
    print("  n_contrib eq", (nc.cpu().numpy().astype(np.uint32) == o["n_contrib"]).mean(), "final_T", rel_l2(fT.cpu().numpy(), o["final_T"]),
          "img", rel_l2(color.cpu().numpy(), o["color"]))
    c = run_cuda(case, dL=dL)
    print("  tight: img", rel_l2(c["color"], o["color"]), *[f"{k} {rel_l2(c[k], o[k]):.2e}" for k in ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D")])

def timing(P_hw=(512, 1024), mode="erp", iters=5):
    H, W = P_hw
    dev = "cuda"
    sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1236, device=dev)
    pose = synthetic.target_pose(3).to(dev)
    cam = camera.erp_camera(pose[None])
    means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
    s = rasterizer.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam.view_matrix[0], projmatrix=cam.full_projection[0], sh_degree=4, campos=cam.campos[0],
        prefiltered=False, debug=False, projection=mode)
    print("P", means.shape[0])
    target = torch.rand(3, H, W, device=dev)
    for it in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record()
        color, st = rasterizer.forward_raw(s, means, cov6, op, shs, None)
        e[1].record()
        g = rasterizer.backward_raw(s, means, cov6, op, shs, None, st, 2 * (color - target) / color.numel())
        e[2].record()
        torch.cuda.synchronize()
        print(f"  it{it}: fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms  N={st.num_rendered} vis={st.num_visible}")
    print("  color mean", float(color.mean()), "finite", bool(torch.isfinite(color).all()), "grad finite", all(bool(torch.isfinite(v).all()) for v in g.values() if v is not None))

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for args in [("pinhole", 96, 128, 3000), ("erp", 64, 128, 3000), ("pinhole", 50, 70, 500), ("erp", 48, 64, 500)]:
        stage_check(*args)
    timing()
