"""Does rendering independent views on two streams raise throughput (latency-bound sorts overlap compute-bound renders)?"""
import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter360_b200 import camera, rasterizer, synthetic
dev = torch.device("cuda", 0)
H, W = 512, 1024
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237, device=dev)
means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
op = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
poses = synthetic.trajectory(32).to(dev); cam = camera.erp_camera(poses)
target = torch.rand(3, H, W, device=dev)
def settings(i):
    return rasterizer.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam.view_matrix[i], projmatrix=cam.full_projection[i], sh_degree=4, campos=cam.campos[i],
        prefiltered=False, debug=False, projection="erp")
def view(i, bwd):
    color, st = rasterizer.forward_raw(settings(i), means, cov6, op, shs, None)
    if bwd:
        rasterizer.backward_raw(settings(i), means, cov6, op, shs, None, st, 2 * (color - target) / color.numel())
def run(nstreams, bwd, n=32):
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(streams[i % nstreams]):
            view(i, bwd)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for bwd in (False, True):
    for ns in (1, 2, 3):
        run(ns, bwd, 8)
        print("bwd" if bwd else "fwd", "streams", ns, "ms/view %.3f" % min(run(ns, bwd) for _ in range(3)))
