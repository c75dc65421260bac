"""Independent float64 PyTorch/autograd restatement -- TEST INFRASTRUCTURE ONLY.

Pure tensor ops, no tiles-as-threads, no hand-written derivatives: the image is composited per
pixel over the depth-sorted Gaussians, and every gradient comes from autograd.  It exists to pin
the hand-derived backward of ``raster_oracle.c`` (and hence the CUDA kernels) on small scenes.

Follows SURVEY.md Appendix A (pinhole) / B2 (erp); call-site conventions from
/root/reference/src/model/decoder/cuda_splatting.py:85-124.  PARITY UNPINNED (see oracle/__init__).

Deliberate non-autograd semantics reproduced with ``detach`` (they are part of the upstream
algorithm, Appendix A K8): when the 1.3*tanfov clamp is active the clamped coordinate is treated
as a constant inside J.  The 3-sigma tile rectangle, the depth order and the clamp masks are
integer/boolean decisions and carry no gradient.
"""
from __future__ import annotations

import math

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]
C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
      -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761]


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """d [P,3] unit directions -> basis [P,(deg+1)^2] (3DGS sign convention)."""
    x, y, z = d.unbind(-1)
    b = [torch.full_like(x, C0)]
    if deg > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    if deg > 1:
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy),
              C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)]
    if deg > 3:
        b += [C4[0] * xy * (xx - yy), C4[1] * yz * (3 * xx - yy), C4[2] * xy * (7 * zz - 1),
              C4[3] * yz * (7 * zz - 3), C4[4] * (zz * (35 * zz - 30) + 3), C4[5] * xz * (7 * zz - 3),
              C4[6] * (xx - yy) * (7 * zz - 1), C4[7] * xz * (xx - 3 * yy),
              C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(b, -1)


def render(means, cov6, opac, *, shs=None, colors=None, H, W, view, proj, campos, bg=(0., 0., 0.),
           tanfovx=1.0, tanfovy=1.0, sh_degree=0, mode="pinhole", near_cull=0.2, fov_clamp=1.3,
           lowpass=0.3, pole_eps=1e-3, max_sh_degree=4, dtype=torch.float64, depth_mode=None, depth_near=0.0,
           depth_far=0.0, depth_scale=1.0):
    """Differentiable render of one view.  Returns (color[3,H,W], aux dict).

    Inputs may require grad.  ``aux['means2D']`` is a dummy leaf whose grad receives the
    NDC-unit screen-space gradient like upstream's ``means2D`` argument.  ``depth_mode``: ``aux['depth']`` [H,W] is the
    differentiable depth channel (per-Gaussian value of sort depth / depth_scale blended with the colour weights,
    the reference's depth-as-colour pass, cuda_splatting.py:226-269).
    """
    means = means.to(dtype)
    cov6 = cov6.to(dtype)
    opac = opac.to(dtype).reshape(-1)
    view = torch.as_tensor(view, dtype=dtype)
    proj = torch.as_tensor(proj, dtype=dtype)
    campos = torch.as_tensor(campos, dtype=dtype)
    bg = torch.as_tensor(bg, dtype=dtype)
    P = means.shape[0]
    R = view[:3, :3].T            # view_i = sum_k R[i,k] world_k
    t = means @ view[:3, :3] + view[3, :3]
    x, y, z = t.unbind(-1)
    S = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4],
                     cov6[:, 2], cov6[:, 4], cov6[:, 5]], -1).reshape(P, 3, 3)
    zero = torch.zeros_like(x)
    means2D = torch.zeros(P, 3, dtype=dtype, requires_grad=True)  # NDC-unit dummy
    if mode == "pinhole":
        keep = z > near_cull
        fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
        limx, limy = fov_clamp * tanfovx, fov_clamp * tanfovy
        zs = torch.where(keep, z, torch.ones_like(z))
        txtz, tytz = x / zs, y / zs
        cx_on = (txtz < -limx) | (txtz > limx)
        cy_on = (tytz < -limy) | (tytz > limy)
        xc = torch.where(cx_on, (txtz.clamp(-limx, limx) * zs).detach(), x)
        yc = torch.where(cy_on, (tytz.clamp(-limy, limy) * zs).detach(), y)
        J = torch.stack([fx / zs, zero, -fx * xc / (zs * zs), zero, fy / zs, -fy * yc / (zs * zs)], -1).reshape(P, 2, 3)
        hom = torch.cat([means, torch.ones(P, 1, dtype=dtype)], -1) @ proj
        pw = 1.0 / (hom[:, 3] + 1e-7)
        ndc = hom[:, :2] * pw[:, None] + means2D[:, :2]
        px = ((ndc[:, 0] + 1) * W - 1) * 0.5
        py = ((ndc[:, 1] + 1) * H - 1) * 0.5
        sortkey = z
    else:
        su, sv = -W / (2 * math.pi), -H / math.pi
        r = t.norm(dim=-1)
        keep = r > near_cull
        rho = torch.sqrt(x * x + z * z)
        pole = rho < pole_eps * r
        s = torch.where(pole & (rho > 0), pole_eps * r / rho.clamp_min(1e-300), torch.ones_like(rho))
        xc = torch.where(pole, (x * s).detach(), x)
        zc = torch.where(pole, torch.where(rho > 0, z * s, pole_eps * r).detach(), z)
        yc = torch.where(pole, y.detach(), y)
        q = xc * xc + zc * zc
        rc = torch.sqrt(q)
        r2 = q + yc * yc
        J = torch.stack([su * zc / q, zero, -su * xc / q,
                         -sv * xc * yc / (rc * r2), sv * rc / r2, -sv * zc * yc / (rc * r2)], -1).reshape(P, 2, 3)
        u = su * torch.atan2(x, z) + 0.5 * W - 0.5
        v = sv * torch.atan2(y, rho) + 0.5 * H - 0.5
        # for pole-clamped Gaussians the position gradient is defined through the clamped J
        tl = t.detach()
        u_lin = u.detach() + (J[:, 0].detach() * (t - tl)).sum(-1)
        v_lin = v.detach() + (J[:, 1].detach() * (t - tl)).sum(-1)
        px = torch.where(pole, u_lin, u) + means2D[:, 0] * (0.5 * W)
        py = torch.where(pole, v_lin, v) + means2D[:, 1] * (0.5 * H)
        sortkey = r
    Mm = J @ R
    cov2 = Mm @ S @ Mm.transpose(1, 2)
    a = cov2[:, 0, 0] + lowpass
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + lowpass
    det = a * c - b * b
    keep = keep & (det != 0)
    dets = torch.where(keep, det, torch.ones_like(det))
    cA, cB, cC = c / dets, -b / dets, a / dets
    # 3-sigma extent -> tile rectangle (integer decisions, float32 like the C oracle)
    a32, b32, c32 = a.detach().float(), b.detach().float(), c.detach().float()
    det32 = a32 * c32 - b32 * b32
    if mode == "pinhole":
        mid = 0.5 * (a32 + c32)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det32, min=0.1))
        ex = ey = torch.ceil(3 * torch.sqrt(lam)).int()
    else:
        ex = torch.ceil(3 * torch.sqrt(a32)).int().clamp(max=W // 2)
        ey = torch.ceil(3 * torch.sqrt(c32)).int()
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pxf, pyf = px.detach().float(), py.detach().float()
    ymin = ((pyf - ey) / 16).int().clamp(0, gy)
    ymax = ((pyf + ey + 15) / 16).int().clamp(0, gy)
    if mode == "pinhole":
        xmin = ((pxf - ex) / 16).int().clamp(0, gx)
        xmax = ((pxf + ex + 15) / 16).int().clamp(0, gx)
    else:
        xmin = torch.floor((pxf - ex) / 16).int()
        xmax = torch.floor((pxf + ex + 15) / 16).int()
        xmax = torch.minimum(xmax, xmin + gx)
    keep = keep & ((xmax - xmin) * (ymax - ymin) > 0)
    # colours
    if shs is not None:
        deg = min(sh_degree, max_sh_degree)
        d = means - campos
        d = d / d.norm(dim=-1, keepdim=True)
        basis = sh_basis(deg, d)
        n = basis.shape[1]
        rgb = (basis[:, :, None] * shs.to(dtype)[:, :n, :]).sum(1) + 0.5
        rgb = torch.clamp_min(rgb, 0.0)
    else:
        rgb = colors.to(dtype)
    # per-pixel compositing in (depth, index) order
    order = sorted([i for i in range(P) if bool(keep[i])], key=lambda i: (float(sortkey[i].detach().float()), i))
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    tyy, txx = (ys / 16).int(), (xs / 16).int()
    dval = None
    if depth_mode is not None:
        zz = sortkey / depth_scale
        eps = 1e-10
        if depth_mode == "disparity":
            dval = 1 / zz
        elif depth_mode == "relative_disparity":
            dn, df = 1 / (depth_near + eps), 1 / (depth_far + eps)
            dval = 1 - (1 / (zz + eps) - df) / (dn - df + eps)
        elif depth_mode == "log":
            dval = zz.minimum(torch.as_tensor(depth_near, dtype=dtype)).maximum(torch.as_tensor(depth_far, dtype=dtype)).log()
        else:
            dval = zz
    D = torch.zeros(H, W, dtype=dtype)
    T = torch.ones(H, W, dtype=dtype)
    C = torch.zeros(3, H, W, dtype=dtype)
    done = torch.zeros(H, W, dtype=torch.bool)
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    count = torch.zeros(H, W, dtype=torch.int64)
    for i in order:
        if mode == "pinhole":
            in_x = (txx >= xmin[i]) & (txx < xmax[i])
        else:
            in_x = torch.remainder(txx - xmin[i], gx) < (xmax[i] - xmin[i])
        member = in_x & (tyy >= ymin[i]) & (tyy < ymax[i])
        count = count + member.long()
        dx = px[i] - xs
        dy = py[i] - ys
        if mode == "erp":
            dx = torch.where(dx > 0.5 * W, dx - W, torch.where(dx < -0.5 * W, dx + W, dx))
        power = -0.5 * (cA[i] * dx * dx + cC[i] * dy * dy) - cB[i] * dx * dy
        alpha = torch.clamp_max(opac[i] * torch.exp(power), 0.99)
        ok = member & (~done) & (power <= 0) & (alpha >= 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = ok & (test_T < 1e-4)
        done = done | stop
        ok = ok & ~stop
        w = torch.where(ok, alpha * T, torch.zeros_like(T))
        C = C + rgb[i][:, None, None] * w
        if dval is not None:
            D = D + dval[i] * w
        T = torch.where(ok, test_T, T)
        n_contrib = torch.where(ok, count, n_contrib)
    color = C + T * bg[:, None, None]
    aux = dict(means2D=means2D, final_T=T.detach(), n_contrib=n_contrib, keep=keep, px=px.detach(), py=py.detach(),
               conic=torch.stack([cA, cB, cC], -1).detach(), rgb=rgb.detach(), ex=ex, ey=ey,
               depth=D if dval is not None else None)
    return color, aux
