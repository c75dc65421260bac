"""Freeze the unknowns of the REAL extension and write golden vectors -- to be run ONCE wherever upstream
`diff_gaussian_rasterization` (dcharatan/diff-gaussian-rasterization-modified, /root/reference/requirements.txt:17) is
importable on a CUDA box, e.g. installed under baseline/_ref:

    python tools/probe_reference_ext.py --ext-path baseline/_ref          # writes tests/golden/ext_probe.npz

Until this has run, every parity number in this repository is "against our restatement of upstream" (DESIGN.md sec. 3).
tests/test_golden.py::test_oracle_matches_the_probed_extension picks the file up as soon as it exists: the oracle must
then reproduce the real extension's images, radii and gradients on the probe set, and the inferred constants
(max_sh_degree, near_cull, fov_clamp, lowpass) must equal oracle.DEFAULTS.

Probes (all tiny, fp32, 64x64 pinhole, tanfov 1):
  sh4       one on-axis Gaussian whose ONLY non-zero SH coefficients are the nine degree-4 ones, seen off-axis:
            colour != 0.5 grey  <=>  the fork evaluates the degree-4 band (max_sh_degree 4); grey <=> stock 3 bands
  near      the same Gaussian moved through z = 0.15 ... 0.25 in 1e-3 steps: first z with radius > 0 = near_cull
  clamp     an off-axis Gaussian at x/z = 1.5 and 2.5 tanfov: cov2D (read back from the rendered footprint's second
            moments) tells whether the centre was clamped to fov_clamp * tanfov inside J
  lowpass   a Gaussian with negligible 3-D extent: footprint variance = lowpass
  recipe    the reference's own smoke recipe (/root/reference/src/scripts/test_splatter.py:38-65): unit-scale Gaussians
            at the origin, degree-2 coefficients = 10, camera on a circle
  cloud     200 random Gaussians, SH degree 4, forward + backward with a fixed dL/dimage: image, radii, all gradients
"""
import argparse
import importlib
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ext(path):
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]   # not this repo's drop-in of the same name
    sys.path.insert(0, os.path.abspath(path))
    m = importlib.import_module("diff_gaussian_rasterization")
    if "splatter360_b200" in sys.modules or not hasattr(m, "_C"):
        raise SystemExit(f"{m.__file__} is not the upstream extension (no _C): point --ext-path at its install directory")
    return m


def camera(H, W, c2w, near=0.01, far=100.0):
    """view / full-projection exactly as /root/reference/src/model/decoder/cuda_splatting.py:80-87 builds them (tanfov 1)."""
    proj = torch.zeros(4, 4)
    proj[0, 0] = 1.0; proj[1, 1] = 1.0; proj[3, 2] = 1.0
    proj[2, 2] = far / (far - near); proj[2, 3] = -(far * near) / (far - near)
    view = c2w.inverse().T
    return view.contiguous(), (view @ proj.T).contiguous(), c2w[:3, 3].contiguous()


def render(ext, dev, H, W, c2w, means, cov6, opac, shs, degree, dL=None):
    view, full, campos = camera(H, W, c2w)
    s = ext.GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
        viewmatrix=view.to(dev), projmatrix=full.to(dev), sh_degree=degree, campos=campos.to(dev), prefiltered=False, debug=False)
    t = [x.to(dev).clone().requires_grad_(dL is not None) for x in (means, cov6, opac[:, None], shs)]
    m2d = torch.zeros_like(t[0], requires_grad=True)
    img, radii = ext.GaussianRasterizer(s)(means3D=t[0], means2D=m2d, shs=t[3], colors_precomp=None, opacities=t[2],
                                            cov3D_precomp=t[1])
    out = dict(image=img.detach().cpu().numpy(), radii=radii.cpu().numpy(), view=view.numpy(), proj=full.numpy(), campos=campos.numpy())
    if dL is not None:
        (img * dL.to(dev)).sum().backward()
        out.update(d_means=t[0].grad.cpu().numpy(), d_cov6=t[1].grad.cpu().numpy(), d_opac=t[2].grad.reshape(-1).cpu().numpy(),
                   d_shs=t[3].grad.cpu().numpy(), d_means2D=m2d.grad.cpu().numpy())
    return out


def moments(img):
    """centroid and second central moments of a single-splat footprint (channel 0)"""
    a = np.clip(img[0].astype(np.float64), 0, None)
    ys, xs = np.mgrid[0:a.shape[0], 0:a.shape[1]]
    w = a.sum()
    cx, cy = (a * xs).sum() / w, (a * ys).sum() / w
    return cx, cy, (a * (xs - cx) ** 2).sum() / w, (a * (ys - cy) ** 2).sum() / w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ext-path", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "ext_probe.npz"))
    args = ap.parse_args()
    ext = load_ext(args.ext_path)
    dev = "cuda"
    H = W = 64
    eye = torch.eye(4)
    out = {}
    iso = lambda s2, n=1: torch.tensor([[s2, 0, 0, s2, 0, s2]], dtype=torch.float32).repeat(n, 1)
    white = torch.zeros(1, 25, 3)

    # sh4: only the degree-4 band is non-zero
    sh = white.clone(); sh[0, 16:25] = 1.0
    r = render(ext, dev, H, W, eye, torch.tensor([[0.7, -0.4, 3.0]]), iso(0.01), torch.tensor([0.9]), sh, 4)
    out["sh4_image"] = r["image"]
    sh0 = white.clone()   # all coefficients zero: colour 0.5 grey
    grey = render(ext, dev, H, W, eye, torch.tensor([[0.7, -0.4, 3.0]]), iso(0.01), torch.tensor([0.9]), sh0, 4)["image"]
    out["max_sh_degree"] = np.array(3 if np.allclose(r["image"], grey, atol=1e-6) else 4)

    # near cull sweep
    zs = np.arange(0.150, 0.2505, 0.001, dtype=np.float32)
    vis = []
    for z in zs:
        rr = render(ext, dev, H, W, eye, torch.tensor([[0.0, 0.0, float(z)]]), iso(1e-6), torch.tensor([0.9]), sh0, 0)
        vis.append(int(rr["radii"][0] > 0))
    out["near_z"] = zs; out["near_visible"] = np.array(vis)
    out["near_cull"] = np.array(float(zs[int(np.argmax(vis))] - 0.001) if any(vis) else np.nan)

    # fov clamp: second moments of an off-axis splat (unclamped J would stretch it more)
    for tag, xz in (("clamp15", 1.5), ("clamp25", 2.5)):
        rr = render(ext, dev, 64, 256, eye, torch.tensor([[xz * 4.0 * 0.55, 0.0, 4.0]]), iso(0.05), torch.tensor([0.9]), sh0, 0)
        out[f"{tag}_image"] = rr["image"]; out[f"{tag}_radii"] = rr["radii"]

    # low-pass: a point-like Gaussian
    rr = render(ext, dev, H, W, eye, torch.tensor([[0.0, 0.0, 4.0]]), iso(1e-8), torch.tensor([0.99]), sh0, 0)
    cx, cy, vx, vy = moments(rr["image"])
    out["lowpass_image"] = rr["image"]; out["lowpass_moments"] = np.array([cx, cy, vx, vy])

    # the reference's smoke recipe
    g = torch.Generator().manual_seed(0)
    n = 3
    A = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))[0]
    cov = A @ A.transpose(1, 2)
    row, col = torch.triu_indices(3, 3)
    shr = torch.zeros(n, 25, 3); shr[:, 4:9, 0] = 10.0
    ang = math.radians(40.0)
    c2w = torch.eye(4); c2w[:3, :3] = torch.tensor([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
    c2w[:3, 3] = c2w[:3, :3] @ torch.tensor([0.0, 0.0, -10.0])
    rr = render(ext, dev, H, W, c2w, torch.zeros(n, 3), cov[:, row, col].contiguous(), torch.ones(n), shr, 4)
    out["recipe_image"] = rr["image"]; out["recipe_radii"] = rr["radii"]; out["recipe_c2w"] = c2w.numpy(); out["recipe_cov6"] = cov[:, row, col].numpy()

    # random cloud, forward + backward
    n = 200
    d = torch.randn(n, 3, generator=g); d = d / d.norm(dim=-1, keepdim=True); d[:, 2] = d[:, 2].abs() + 0.3
    depth = 1.0 + 4.0 * torch.rand(n, generator=g)
    means = d * depth[:, None]
    B = torch.randn(n, 3, 3, generator=g) * (0.05 * depth)[:, None, None]
    cov = B @ B.transpose(1, 2) + 1e-4 * torch.eye(3)
    shc = torch.randn(n, 25, 3, generator=g) * 0.3; shc[:, 0] += 1.0
    opac = 0.1 + 0.85 * torch.rand(n, generator=g)
    dL = torch.randn(3, H, W, generator=g)
    rr = render(ext, dev, H, W, eye, means, cov[:, row, col].contiguous(), opac, shc, 4, dL=dL)
    for k, v in rr.items():
        out[f"cloud_{k}"] = v
    out.update(cloud_means=means.numpy(), cloud_cov6=cov[:, row, col].numpy(), cloud_opac=opac.numpy(), cloud_shs=shc.numpy(),
               cloud_dL=dL.numpy())
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "max_sh_degree", int(out["max_sh_degree"]), "near_cull", float(out["near_cull"]))


if __name__ == "__main__":
    main()
