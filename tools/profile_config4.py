"""Host-side profile of the config-4 step (DecoderSplattingCUDA six faces + Cube2Equirec + MSE, fwd+bwd): cProfile over 30
steps (no syncs inside), top entries by cumulative time; plus wall-clock per step with and without a device sync."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from splatter360_b200 import cubemap, synthetic
from splatter360_b200.decoder import DecoderSplattingCUDA, Gaussians
from splatter360_b200.loss import mse_loss

dev = "cuda"
H, W, Fw = 512, 1024, 256
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1238, device=dev)
g = Gaussians(sc.means[None].contiguous().requires_grad_(), sc.covariances[None].contiguous().requires_grad_(),
              sc.harmonics[None].contiguous().requires_grad_(), sc.opacities[None].contiguous().requires_grad_())
poses = synthetic.trajectory(64, seed=0).to(dev)
faces = cubemap.cube_face_extrinsics(poses)
Kf = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev).expand(1, 6, 3, 3)
near = torch.ones(1, 6, device=dev); far = torch.full((1, 6), 100.0, device=dev)
dec = DecoderSplattingCUDA(sync_free=True).to(dev)
c2e = cubemap.Cube2Equirec(Fw, H, W).to(dev)
target = torch.rand(1, 3, H, W, device=dev)

def step(i):
    for t in (g.means, g.covariances, g.harmonics, g.opacities):
        t.grad = None
    out = dec(g, faces[i][None], Kf, near, far, (Fw, Fw))
    pano = c2e.from_faces(out.color)
    loss = mse_loss(pano, target)
    loss.backward()
    return loss

for i in range(5):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(30):
    step(5 + i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per step {1e3 * (t1 - t0) / 30:.3f} ms; wall per step incl. final sync {1e3 * (t2 - t0) / 30:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for i in range(30):
    step(5 + i)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])

# device side: every kernel / memset / memcpy of the step by total time (torch.profiler, 10 steps)
try:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(10):
            step(5 + i)
        torch.cuda.synchronize()
    rows = sorted(((e.key, e.device_time_total / 10, e.count / 10) for e in prof.key_averages() if e.device_time_total > 0),
                  key=lambda r: -r[1])
    print(f"device time per step by kernel (us), total {sum(r[1] for r in rows):.1f} us over {sum(r[2] for r in rows):.0f} launches")
    for k, us, n in rows[:45]:
        print(f"{us:9.1f} us  x{n:4.1f}  {k[:110]}")
except Exception as e:  # CUPTI may be unavailable on the box
    print("torch.profiler unavailable:", e)
