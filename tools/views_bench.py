"""Batched six-face pass at BASELINE.json configs[2] scale (1,048,576 pixel-aligned Gaussians, six 256x256 faces =
the reference's 512x1024 panorama): CUDA-event totals + the library's per-stage timers, cameras precomputed.
S360_LIB selects a variant build (tools/ab_views.sh)."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter360_b200 import camera, cubemap, rasterizer, synthetic, _lib

dev = "cuda"
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc = synthetic.pixel_aligned_scene(2 * F, 4 * F, seed=1237, device=dev)
means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
pose = synthetic.trajectory(1, seed=1).to(dev)[0]
faces = cubemap.cube_face_extrinsics(pose)
K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None].repeat(6, 1, 1)
cam = camera.pinhole_camera(faces, K, torch.ones(6, device=dev), torch.full((6,), 100.0, device=dev))
mk = lambda vm, pm, cp: rasterizer.GaussianRasterizationSettings(
    image_height=F, image_width=F, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
    viewmatrix=vm.contiguous(), projmatrix=pm.contiguous(), sh_degree=4, campos=cp.contiguous(), prefiltered=False, debug=False,
    projection="pinhole")
s6 = mk(cam.view_matrix, cam.full_projection, cam.campos)
s1 = [mk(cam.view_matrix[k], cam.full_projection[k], cam.campos[k]) for k in range(6)]
dL = torch.rand(6, 3, F, F, device=dev) / (3 * F * F)
info = {}

def batched(bwd):
    color, st = rasterizer.forward_views_raw(s6, means, cov6, op, shs, None)
    info.update(pairs=st.num_pairs, N=st.num_rendered, pair_capacity=st.pair_capacity)
    if bwd:
        rasterizer.backward_views_raw(s6, means, cov6, op, shs, None, st, dL)

def separate(bwd):
    for k in range(6):
        color, st = rasterizer.forward_raw(s1[k], means, cov6, op, shs, None)
        if bwd:
            rasterizer.backward_raw(s1[k], means, cov6, op, shs, None, st, dL[k])

def timeit(fn, iters=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)

if "--profile" in sys.argv:   # under ncu: a few forward+backward passes, nothing else
    for _ in range(4):
        batched(True)
    torch.cuda.synchronize()
    sys.exit(0)
res = {"lib": os.environ.get("S360_LIB", "default"), "face": F, "P": means.shape[0]}
res["batched_fwd_ms"] = timeit(lambda: batched(False))
res["batched_fwd_bwd_ms"] = timeit(lambda: batched(True))
if "--no-separate" not in sys.argv:
    res["separate_fwd_ms"] = timeit(lambda: separate(False), iters=10, warm=2)
    res["separate_fwd_bwd_ms"] = timeit(lambda: separate(True), iters=10, warm=2)
_lib.profile_enable(True); _lib.profile_read(reset=True)
for _ in range(10): batched(True)
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True); _lib.profile_enable(False)
res["stages_ms"] = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()}
res["stage_sum_ms"] = round(sum(res["stages_ms"].values()), 4)
res.update(info)
print(json.dumps(res)); sys.stdout.flush()
