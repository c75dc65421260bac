import sys, os, time, json, torch
sys.path.insert(0, os.getcwd())
from splatter360_b200 import synthetic, camera
from splatter360_b200.decoder import render_erp
from splatter360_b200.loss import mse_loss
from splatter360_b200.io import HostSceneFeeder
dev = torch.device("cuda:0")
H, W = 512, 1024
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237, device=dev)
host = dict(means=sc.means.cpu().pin_memory(), cov=sc.covariances.cpu().pin_memory(), sh=sc.harmonics.cpu().pin_memory(),
            op=sc.opacities.cpu().pin_memory(), target=torch.rand(3, H, W).pin_memory())
nbytes = sum(t.numel() * 4 for t in host.values())
# raw H2D
dst = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
for _ in range(2):
    for k in host: dst[k].copy_(host[k], non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    for k in host: dst[k].copy_(host[k], non_blocking=True)
b.record(); torch.cuda.synchronize()
print("raw H2D GB/s", nbytes * 5 / (a.elapsed_time(b) * 1e-3) / 1e9)
poses = synthetic.trajectory(30, seed=0)
hp = poses.pin_memory()
near = torch.ones(1, device=dev); far = torch.full((1,), 100.0, device=dev); bg = torch.zeros(1, 3, device=dev)
feeder = HostSceneFeeder(dev)
def inputs(i): return dict(host, pose=hp[i:i + 1])
def compute(d):
    m, c, sh, o = (d[k].requires_grad_() for k in ("means", "cov", "sh", "op"))
    img = render_erp(d["pose"], near, far, (H, W), bg, m[None], c[None], sh[None], o[None], scale_invariant=False)
    loss = mse_loss(img[0], d["target"]); loss.backward()
    return float(loss.item())
def run(n):
    t = feeder.submit(inputs(0))
    for i in range(n):
        d = feeder.get(t)
        if i + 1 < n: t = feeder.submit(inputs(i + 1))
        compute(d)
run(3); torch.cuda.synchronize()
s0 = torch.cuda.memory_stats()
t0 = time.perf_counter(); a.record(); run(20); b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
s1 = torch.cuda.memory_stats()
print("e2e ms/step", a.elapsed_time(b) / 20, "wall", (t1 - t0) / 20 * 1e3)
for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "segment.all.allocated", "reserved_bytes.all.peak"):
    print(k, s0.get(k), "->", s1.get(k))
# compute only (device-resident), same API
d = {k: v.to(dev) for k, v in inputs(0).items()}
for _ in range(3): compute(d)
torch.cuda.synchronize(); a.record()
for _ in range(20): compute(d)
b.record(); torch.cuda.synchronize()
print("compute-only ms/step", a.elapsed_time(b) / 20)
# upload only through the feeder
torch.cuda.synchronize(); a.record()
for i in range(10):
    t = feeder.submit(inputs(i)); feeder.get(t)
b.record(); torch.cuda.synchronize()
print("feeder upload-only ms/step", a.elapsed_time(b) / 10)
os.system("nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader; nproc; cat /proc/meminfo | head -2")
