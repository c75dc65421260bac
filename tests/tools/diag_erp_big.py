import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_case, run_cuda, run_oracle, rel_l2
case = make_case(40000, "erp", 1024, 2048, seed=7)
dL = torch.randn(3, 1024, 2048, generator=torch.Generator().manual_seed(1))
o = run_oracle(case, dL=dL); c = run_cuda(case, dL=dL)
for k in ("color", "d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"):
    print(k, rel_l2(c[k], o[k]))
err = np.linalg.norm(c["d_means"] - o["d_means"], axis=1); ref = np.linalg.norm(o["d_means"], axis=1)
idx = np.argsort(-err)[:12]
tot = (err ** 2).sum()
V = case["view"].numpy()
for i in idx:
    m = case["means"][i].numpy(); t = m @ V[:3, :3] + V[3, :3]
    rho = np.hypot(t[0], t[2]); r = np.linalg.norm(t)
    print(i, "err", err[i], "ref", ref[i], "share", err[i] ** 2 / tot, "lat_deg", np.degrees(np.arctan2(t[1], rho)), "rho/r", rho / r,
          "xy", o["xy"][i], "m2d err", np.abs(c["d_means2D"][i] - o["d_means2D"][i]).max(), "m2d", np.abs(o["d_means2D"][i]).max())
n = 40000
for k in ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"):
    a = c[k].reshape(n, -1).astype(np.float64); b = o[k].reshape(n, -1).astype(np.float64)
    per = np.linalg.norm(a - b, axis=1) / (np.linalg.norm(b, axis=1) + 1e-12 * np.linalg.norm(b))
    keep = per <= 1e-3
    print(k, "frac>1e-3", np.mean(per > 1e-3), "frac>1e-2", np.mean(per > 1e-2), "frac>1e-4", np.mean(per > 1e-4), "rel_l2 kept", rel_l2(a[keep], b[keep]), "median", np.median(per))
