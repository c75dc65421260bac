#!/bin/bash
# compute-sanitizer over the small parity cases (SURVEY.md sec. 4b tooling row)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from helpers import make_case, run_cuda
for mode, H, W, n in (("pinhole", 64, 80, 1500), ("erp", 48, 96, 1500), ("erp", 37, 64, 300), ("erp", 64, 128, 4500)):   # 4500: three binning chunks
    case = make_case(n, mode, H, W, seed=7)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1))
    out = run_cuda(case, dL=dL)
    case["colors"] = torch.rand(n, 3)
    out = run_cuda(case, dL=dL, use_sh=False)
    print(mode, "ok", float(out["color"].mean()))
# batched multi-view pass (six cube faces, three erp views) + the stitch kernel
import test_gpu_views as tv
from splatter360_b200 import cubemap
sc = tv._scene(1200, seed=3)
for mode, cam, H, W in (("pinhole", tv._cube_cameras(4), 48, 48), ("erp", tv._erp_cameras(3, seed=1), 32, 64)):
    V = cam.view_matrix.shape[0]
    dL = torch.randn(V, 3, H, W, generator=torch.Generator().manual_seed(2))
    out = tv._run_views(sc, tv._settings(cam, H, W, mode, "cuda"), dL=dL)
    print("batched", mode, "ok", float(out["color"].mean()))
c2e = cubemap.Cube2Equirec(16, 32, 64).to("cuda")
f = torch.rand(1, 6, 3, 16, 16, device="cuda", requires_grad=True)
c2e.from_faces(f).sum().backward()
print("stitch ok", float(f.grad.sum()))
# fused Gaussian adapter (ragged size: the last CTA takes the non-bulk path)
from splatter360_b200 import adapter
mod = adapter.GaussianAdapterERP(adapter.GaussianAdapterERPCfg(0.5, 15.0, 4)).to("cuda")
b, v, h, w = 1, 2, 17, 23
ext = torch.eye(4, device="cuda").repeat(b, v, 1, 1)[:, :, None, None, None]
dep = (0.5 + torch.rand(b, v, h * w, 1, 1, device="cuda")).requires_grad_()
raw = torch.randn(b, v, h * w, 1, 1, 82, device="cuda").requires_grad_()
g = mod("hm3d", ext, dep, torch.ones(b, v, h * w, 1, 1, device="cuda"), raw, (h, w))
(g.means.sum() + g.covariances.sum() + g.harmonics.sum()).backward()
print("adapter ok", float(raw.grad.abs().sum()))
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout -s KILL 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python /tmp/san_case.py > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|error|hazard" gpurun_out/r02_sanitize_$tool.log | head -8
done
