"""How much host time does one forward+backward cost?  (The GPU needs ~1.2 ms; the host must stay below that.)"""
import cProfile, pstats, sys, os, time, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter360_b200 import camera, synthetic
from splatter360_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
dev = torch.device("cuda", 0)
H, W = 512, 1024
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237, device=dev)
means = sc.means.contiguous().requires_grad_(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous().requires_grad_()
opac = sc.opacities[:, None].contiguous().requires_grad_(); shs = sc.harmonics.permute(0, 2, 1).contiguous().requires_grad_()
poses = synthetic.trajectory(64).to(dev); cams = camera.erp_camera(poses)
bg = torch.zeros(3, device=dev); target = torch.rand(3, H, W, device=dev)
def step(i):
    s = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0,
        viewmatrix=cams.view_matrix[i], projmatrix=cams.full_projection[i], sh_degree=4, campos=cams.campos[i], prefiltered=False, debug=False, projection="erp")
    m2d = torch.zeros_like(means, requires_grad=True)
    for t in (means, cov6, opac, shs): t.grad = None
    color, _ = GaussianRasterizer(s)(means3D=means, means2D=m2d, shs=shs, colors_precomp=None, opacities=opac, cov3D_precomp=cov6)
    loss = ((color - target) ** 2).mean(); loss.backward()
for i in range(10): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50): step(i % 64)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue time per step {1e3*(t1-t0)/50:.3f} ms ; wall incl. drain {1e3*(t2-t0)/50:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for i in range(30): step(i % 64)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
