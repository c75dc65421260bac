"""Shared test helpers: seeded scenes, cameras, oracle/CUDA drivers, error norms."""
from __future__ import annotations

import numpy as np
import torch

from splatter360_b200 import camera, synthetic


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_case(n: int, mode: str, H: int, W: int, seed: int = 0, sh_degree: int = 4, inflate: float | None = None,
              depth_range=(0.5, 4.0)):
    """Random-cloud scene + one camera.  Returns dict of CPU tensors ready for oracle and CUDA paths."""
    sc = synthetic.random_cloud_scene(n, sh_degree=sh_degree, seed=seed, ref_width=1024, depth_range=depth_range)
    if inflate is None:
        inflate = 1024.0 / max(H, W)
    cov = sc.covariances * inflate ** 2
    pose = synthetic.target_pose(seed)
    if mode == "pinhole":
        K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None]
        cam = camera.pinhole_camera(pose[None], K, torch.tensor([1.0]), torch.tensor([100.0]))
    else:
        cam = camera.erp_camera(pose[None])
    return dict(
        means=sc.means.contiguous(),
        cov6=synthetic.cov3x3_to_cov6(cov).contiguous(),
        opac=sc.opacities.contiguous(),
        shs=sc.harmonics.permute(0, 2, 1).contiguous(),
        H=H, W=W, mode=mode, sh_degree=sh_degree,
        view=cam.view_matrix[0].contiguous(), proj=cam.full_projection[0].contiguous(),
        campos=cam.campos[0].contiguous(), tanfovx=float(cam.tan_fov_x[0]), tanfovy=float(cam.tan_fov_y[0]),
        bg=torch.tensor([0.1, 0.2, 0.3]),
    )


def oracle_kwargs(case, **over):
    kw = dict(H=case["H"], W=case["W"], view=case["view"].numpy(), proj=case["proj"].numpy(),
              campos=case["campos"].numpy(), bg=case["bg"].numpy(), tanfovx=case["tanfovx"],
              tanfovy=case["tanfovy"], sh_degree=case["sh_degree"], mode=case["mode"])
    kw.update(over)
    return kw


def run_oracle(case, dL=None, use_sh=True, stages=True, **over):
    import oracle
    kw = oracle_kwargs(case, **over)
    if use_sh:
        return oracle.render(case["means"].numpy(), case["cov6"].numpy(), case["opac"].numpy(),
                             shs=case["shs"].numpy(), dL_dpix=None if dL is None else dL.numpy(), stages=stages, **kw)
    return oracle.render(case["means"].numpy(), case["cov6"].numpy(), case["opac"].numpy(),
                         colors=case["colors"].numpy(), dL_dpix=None if dL is None else dL.numpy(), stages=stages, **kw)


def make_settings(case, device="cuda", **over):
    from splatter360_b200.rasterizer import GaussianRasterizationSettings
    kw = dict(
        image_height=case["H"], image_width=case["W"], tanfovx=case["tanfovx"], tanfovy=case["tanfovy"],
        bg=case["bg"].to(device), scale_modifier=1.0, viewmatrix=case["view"].to(device),
        projmatrix=case["proj"].to(device), sh_degree=case["sh_degree"], campos=case["campos"].to(device),
        prefiltered=False, debug=False, projection=case["mode"])
    kw.update(over)
    return GaussianRasterizationSettings(**kw)


def run_cuda(case, dL=None, use_sh=True, device="cuda", **over):
    """Forward (+backward) through the public GaussianRasterizer.  Returns dict of CPU numpy arrays."""
    from splatter360_b200.rasterizer import GaussianRasterizer
    settings = make_settings(case, device, **over)
    means = case["means"].to(device).requires_grad_()
    cov6 = case["cov6"].to(device).requires_grad_()
    opac = case["opac"].to(device)[:, None].clone().requires_grad_()
    m2d = torch.zeros_like(means, requires_grad=True)
    if use_sh:
        feat = case["shs"].to(device).requires_grad_()
        color, radii = GaussianRasterizer(settings)(means3D=means, means2D=m2d, shs=feat, colors_precomp=None,
                                                    opacities=opac, cov3D_precomp=cov6)
    else:
        feat = case["colors"].to(device).requires_grad_()
        color, radii = GaussianRasterizer(settings)(means3D=means, means2D=m2d, shs=None, colors_precomp=feat,
                                                    opacities=opac, cov3D_precomp=cov6)
    out = dict(color=color.detach().cpu().numpy(), radii=radii.cpu().numpy())
    if dL is not None:
        (color * dL.to(device)).sum().backward()
        out.update(d_means=means.grad.cpu().numpy(), d_cov6=cov6.grad.cpu().numpy(),
                   d_opac=opac.grad.reshape(-1).cpu().numpy(), d_means2D=m2d.grad.cpu().numpy())
        out["d_shs" if use_sh else "d_colors"] = feat.grad.cpu().numpy()
    return out
