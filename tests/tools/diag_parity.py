"""Where does the GPU-vs-oracle difference at config 3 come from?  Prints rel-L2 per output for an image-like and a
white-noise seed gradient, split by latitude band (poles vs equator) and by splat radius, plus the worst Gaussians."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, run_cuda, run_oracle
import test_gpu_baseline_configs as B
from splatter360_b200 import synthetic

H, W = 512, 1024
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237)
case = B._erp_case(sc, H, W, synthetic.trajectory(8, seed=0)[3])
out = {}
for tag, dL in (("image_like", B.smooth_seed(H, W, 3)), ("white", torch.randn(3, H, W, generator=torch.Generator().manual_seed(3)) / (3 * H * W))):
    o = run_oracle(case, dL=dL, stages=True)
    c = run_cuda(case, dL=dL)
    r = {k: rel_l2(c[k], o[k]) for k in ("color", "d_means", "d_cov6", "d_opac", "d_shs", "d_means2D")}
    y = o["xy"][:, 1]
    vis = o["radii"] > 0
    bands = {"polar(|v-H/2|>0.4H)": vis & (np.abs(y - H / 2) > 0.4 * H), "mid": vis & (np.abs(y - H / 2) <= 0.4 * H) & (np.abs(y - H / 2) > 0.2 * H),
             "equator": vis & (np.abs(y - H / 2) <= 0.2 * H)}
    r["by_latitude"] = {b: {k: rel_l2(c[k][m], o[k][m]) for k in ("d_means", "d_cov6")} for b, m in bands.items()}
    rad = o["radii"]
    r["by_radius"] = {f"r<={hi}": {k: rel_l2(c[k][vis & (rad <= hi) & (rad > lo)], o[k][vis & (rad <= hi) & (rad > lo)]) for k in ("d_means", "d_cov6")}
                      for lo, hi in ((0, 3), (3, 6), (6, 12), (12, 10000))}
    err = np.linalg.norm(c["d_cov6"].astype(np.float64) - o["d_cov6"], axis=1)
    top = np.argsort(-err)[:8]
    r["worst_d_cov"] = [dict(i=int(i), err=float(err[i]), norm=float(np.linalg.norm(o["d_cov6"][i])), radius=int(rad[i]), y=float(y[i]),
                             x=float(o["xy"][i, 0]), opac=float(case["opac"][i]), depth=float(o["depth"][i])) for i in top]
    r["share_of_sq_error_in_top_1000"] = float((np.sort(err ** 2)[-1000:]).sum() / (err ** 2).sum())
    r["radii_mismatch"] = int((c["radii"] != o["radii"]).sum())
    r["xy_max_abs_diff_vs_forward_stage"] = None
    out[tag] = r
print(json.dumps(out, indent=1))
