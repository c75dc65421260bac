"""Host vs device time for the reference-style six-face loop at 1M Gaussians (pinhole 256x256 faces)."""
import cProfile, pstats, sys, os, time, io, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter360_b200 import camera, rasterizer, synthetic, cubemap, _lib
dev = torch.device("cuda", 0)
H, W = 512, 1024
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237, device=dev)
means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
pose = synthetic.target_pose(3).to(dev)
fp = cubemap.cube_face_extrinsics(pose)
K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None].repeat(6, 1, 1)
camp = camera.pinhole_camera(fp, K, torch.ones(6, device=dev), torch.full((6,), 100.0, device=dev))
def settings(f):
    return rasterizer.GaussianRasterizationSettings(image_height=256, image_width=256, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=camp.view_matrix[f], projmatrix=camp.full_projection[f], sh_degree=4, campos=camp.campos[f],
        prefiltered=False, debug=False, projection="pinhole")
def pano():
    for f in range(6):
        rasterizer.forward_raw(settings(f), means, cov6, op, shs, None)
for _ in range(5): pano()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): pano()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"6 faces: host {1e3*(t1-t0)/20:.3f} ms, wall {1e3*(t2-t0)/20:.3f} ms")
_lib.profile_read(True); _lib.profile_enable(True)
for _ in range(10): pano()
torch.cuda.synchronize(); _lib.profile_enable(False)
st = _lib.profile_read(True)
print({k: round(v[0] / max(v[1], 1), 4) for k, v in st.items()}, "sum per face", round(sum(v[0] / max(v[1], 1) for v in st.values()), 4))
pr = cProfile.Profile(); pr.enable()
for _ in range(10): pano()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(s.getvalue()[:2400])
