"""Measure every BASELINE.json config on one GPU (CUDA events, median of 30 after 5 warm-ups) and the CPU oracle on
config 1.  Writes gpurun_out/configs.json; the table goes into BASELINE.md / profiles/."""
import json, math, os, sys, time, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from splatter360_b200 import camera, rasterizer, synthetic, _lib

dev = "cuda"

def face_poses(c2w):
    def rx(a):
        c, s = math.cos(a), math.sin(a); return torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float32)
    def ry(a):
        c, s = math.cos(a), math.sin(a); return torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float32)
    out = []
    for R in (rx(math.pi / 2), torch.eye(3), ry(-math.pi / 2), ry(-math.pi), ry(-1.5 * math.pi), rx(-math.pi / 2)):
        m = c2w.clone(); m[:3, :3] = c2w[:3, :3] @ R.to(c2w.device); out.append(m)
    return torch.stack(out)

def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return dict(median_ms=statistics.median(ts), p10_ms=ts[len(ts) // 10], p90_ms=ts[(len(ts) * 9) // 10])

def settings(H, W, cam, i, mode, tan=1.0):
    return rasterizer.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=tan, tanfovy=tan, bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam.view_matrix[i], projmatrix=cam.full_projection[i], sh_degree=4, campos=cam.campos[i],
        prefiltered=False, debug=False, projection=mode)

def run(name, scene, H, W, n_views, backward):
    means = scene.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(scene.covariances).contiguous()
    op = scene.opacities.contiguous(); shs = scene.harmonics.permute(0, 2, 1).contiguous()
    P = means.shape[0]
    poses = synthetic.trajectory(n_views, seed=1).to(dev)
    cam = camera.erp_camera(poses)
    target = torch.rand(3, H, W, device=dev)
    info = {}
    def erp_step(bwd):
        for i in range(n_views):
            color, st = rasterizer.forward_raw(settings(H, W, cam, i, "erp"), means, cov6, op, shs, None)
            info.update(N=st.num_rendered, P_vis=st.num_visible)
            if bwd:
                rasterizer.backward_raw(settings(H, W, cam, i, "erp"), means, cov6, op, shs, None, st, 2 * (color - target) / color.numel())
    res = {"config": name, "P": P, "image": [H, W], "views": n_views}
    res["erp_fwd"] = timeit(lambda: erp_step(False))
    if backward:
        res["erp_fwd_bwd"] = timeit(lambda: erp_step(True))
    res.update(info)
    # reference-style panorama: 6 pinhole cube faces of edge H/2 (model_wrapper_erp.py:202-205), no stitching cost
    F = H // 2
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None].repeat(6, 1, 1)
    tgt6 = torch.rand(3, F, F, device=dev)
    camps = [camera.pinhole_camera(face_poses(poses[i]), K, torch.ones(6, device=dev), torch.full((6,), 100.0, device=dev)) for i in range(n_views)]
    def cube_step(bwd):
        for i in range(n_views):
            camp = camps[i]
            for f in range(6):
                color, st = rasterizer.forward_raw(settings(F, F, camp, f, "pinhole"), means, cov6, op, shs, None)
                if bwd:
                    rasterizer.backward_raw(settings(F, F, camp, f, "pinhole"), means, cov6, op, shs, None, st, 2 * (color - tgt6) / color.numel())
    res["cube6_fwd"] = timeit(lambda: cube_step(False), iters=10, warm=2)
    if backward:
        res["cube6_fwd_bwd"] = timeit(lambda: cube_step(True), iters=10, warm=2)
    # the same six faces in ONE batched pass (s360_multi_*), plus the stitch kernel (change_order + Cube2Equirec)
    from splatter360_b200 import cubemap
    c2e = cubemap.Cube2Equirec(F, H, W).to(dev)
    tgt_pano = torch.rand(3, H, W, device=dev)
    stats = {}
    def cube_batched_step(bwd, stitch):
        for i in range(n_views):
            camp = camps[i]
            s6 = rasterizer.GaussianRasterizationSettings(image_height=F, image_width=F, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev),
                scale_modifier=1.0, viewmatrix=camp.view_matrix, projmatrix=camp.full_projection, sh_degree=4, campos=camp.campos,
                prefiltered=False, debug=False, projection="pinhole")
            color, st = rasterizer.forward_views_raw(s6, means, cov6, op, shs, None)
            stats.update(pairs=st.num_pairs, N6=st.num_rendered)
            if stitch:
                faces = color[None].detach().requires_grad_(bwd)
                pano = c2e.from_faces(faces)
                if bwd:
                    (g_faces,) = torch.autograd.grad(pano, faces, 2 * (pano[0] - tgt_pano)[None] / pano.numel())
                    rasterizer.backward_views_raw(s6, means, cov6, op, shs, None, st, g_faces[0])
            elif bwd:
                rasterizer.backward_views_raw(s6, means, cov6, op, shs, None, st, 2 * (color - tgt6) / color.numel())
    res["cube6_batched_fwd"] = timeit(lambda: cube_batched_step(False, False), iters=20, warm=3)
    res["cube6_batched_stitched_fwd"] = timeit(lambda: cube_batched_step(False, True), iters=20, warm=3)
    if backward:
        res["cube6_batched_fwd_bwd"] = timeit(lambda: cube_batched_step(True, False), iters=20, warm=3)
        res["cube6_batched_stitched_fwd_bwd"] = timeit(lambda: cube_batched_step(True, True), iters=20, warm=3)
    res.update(stats)
    for k in ("erp_fwd", "erp_fwd_bwd", "cube6_fwd", "cube6_fwd_bwd", "cube6_batched_fwd", "cube6_batched_fwd_bwd",
              "cube6_batched_stitched_fwd", "cube6_batched_stitched_fwd_bwd"):
        if k in res:
            res[k]["views_per_s"] = n_views / (res[k]["median_ms"] * 1e-3)
            res[k]["gaussians_per_s"] = P * n_views / (res[k]["median_ms"] * 1e-3)
    print(json.dumps(res)); sys.stdout.flush()
    return res

def stage_profile(scene, F):
    """Per-stage CUDA-event times of the batched six-face pass (library stage timers)."""
    means = scene.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(scene.covariances).contiguous()
    op = scene.opacities.contiguous(); shs = scene.harmonics.permute(0, 2, 1).contiguous()
    poses = synthetic.trajectory(1, seed=1).to(dev)
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None].repeat(6, 1, 1)
    camp = camera.pinhole_camera(face_poses(poses[0]), K, torch.ones(6, device=dev), torch.full((6,), 100.0, device=dev))
    s6 = rasterizer.GaussianRasterizationSettings(image_height=F, image_width=F, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=camp.view_matrix, projmatrix=camp.full_projection, sh_degree=4, campos=camp.campos,
        prefiltered=False, debug=False, projection="pinhole")
    dL = torch.rand(6, 3, F, F, device=dev)
    def step():
        color, st = rasterizer.forward_views_raw(s6, means, cov6, op, shs, None)
        rasterizer.backward_views_raw(s6, means, cov6, op, shs, None, st, dL)
        return st
    for _ in range(3): step()
    torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_read(reset=True)
    for _ in range(10): st = step()
    torch.cuda.synchronize()
    prof = _lib.profile_read(reset=True); _lib.profile_enable(False)
    r = {"config": "batched six-face pass, per-stage ms (CUDA events, mean of 10)", "pairs": st.num_pairs, "N": st.num_rendered,
         "stages_ms": {k: v[0] / max(v[1], 1) for k, v in prof.items()}}
    print(json.dumps(r)); sys.stdout.flush()
    return r

out = []
out.append(run("1: 10k random cloud, 256x512 ERP", synthetic.random_cloud_scene(10000, seed=1235, device=dev), 256, 512, 1, True))
out.append(run("2: 300k random cloud, 512x1024 ERP, forward", synthetic.random_cloud_scene(300000, seed=1236, device=dev), 512, 1024, 1, True))
out.append(run("3: 1,048,576 pixel-aligned, 512x1024 ERP, fwd+bwd", synthetic.pixel_aligned_scene(512, 1024, seed=1237, device=dev), 512, 1024, 1, True))
out.append(stage_profile(synthetic.pixel_aligned_scene(512, 1024, seed=1237, device=dev), 256))
if "--skip5" not in sys.argv:
    out.append(run("5: 3M random cloud, 1024x2048 ERP, 4 frames/GPU, forward", synthetic.random_cloud_scene(3000000, seed=1239, ref_width=2048, device=dev), 1024, 2048, 4, False))
# config 1 on the CPU oracle (the 'CPU PyTorch alpha-composite reference' slot: the reference has no CPU path)
import numpy as np, oracle
sc = synthetic.random_cloud_scene(10000, seed=1235)
cam = camera.erp_camera(synthetic.trajectory(1, seed=1))
args = (sc.means.numpy(), synthetic.cov3x3_to_cov6(sc.covariances).numpy(), sc.opacities.numpy())
kw = dict(shs=sc.harmonics.permute(0, 2, 1).contiguous().numpy(), H=256, W=512, view=cam.view_matrix[0].numpy(), proj=cam.full_projection[0].numpy(),
          campos=cam.campos[0].numpy(), sh_degree=4, mode="erp", stages=False)
dL = np.random.default_rng(0).standard_normal((3, 256, 512)).astype(np.float32)
t = []
for _ in range(5):
    t0 = time.perf_counter(); oracle.render(*args, **kw); t1 = time.perf_counter(); oracle.render(*args, dL_dpix=dL, **kw); t2 = time.perf_counter()
    t.append((t1 - t0, t2 - t1))
cpu = {"config": "1 (CPU oracle, C + OpenMP)", "threads": oracle.num_threads(), "fwd_s": min(x[0] for x in t), "fwd_bwd_s": min(x[1] for x in t)}
print(json.dumps(cpu))
out.append(cpu)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
