"""SURVEY.md sec. 4b "drop-in integration": a loop shaped like the reference decoder's forward
(/root/reference/src/model/decoder/decoder_splatting_cuda.py:34-70 -> cuda_splatting.py:47-127: b = 1, v = 18 pinhole views of
256x256, one scene of 1,048,576 Gaussians, one `GaussianRasterizer` call per view with the STOCK upstream arguments -- 12-field
settings, SH as [G, 25, 3], covariances as the six upper-triangle values, pre-scaled copies of the scene) running on the module
named `diff_gaussian_rasterization`, i.e. this library found by the name the reference imports.  Needs no reference checkout.

Checked: the 18 images equal what `DecoderSplattingCUDA` (one batched pass per six faces, folded layouts, scale on load)
produces for the same inputs, two of them equal the CPU oracle, and the gradients of a loss over all 18 images agree between
the per-view loop (autograd sums 18 backward passes) and the batched decoder."""
from math import isqrt

import numpy as np
import pytest
import torch

from helpers import rel_l2, run_oracle

pytestmark = pytest.mark.gpu


def _stock_view_loop(extrinsics, intrinsics, near, far, hw, background, means, covariances, harmonics, opacities):
    """One batch item, all views, the way the reference issues them: per view a rescaled copy of the scene (1 / near), the
    camera matrices in the row-vector convention and one call through the upstream Python API."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer   # the drop-in module name
    from splatter360_b200 import camera
    h, w = hw
    degree = isqrt(harmonics.shape[-1]) - 1
    upper = torch.triu_indices(3, 3)
    sh_upstream = harmonics.transpose(-1, -2).contiguous()                       # [G, 3, 25] -> [G, 25, 3]
    images = []
    for v in range(extrinsics.shape[0]):
        s = 1.0 / near[v]
        c2w = extrinsics[v].clone()
        c2w[:3, 3] = c2w[:3, 3] * s
        fov = camera.get_fov(intrinsics[v][None])[0]
        proj = camera.get_projection_matrix((near[v] * s)[None], (far[v] * s)[None], fov[0][None], fov[1][None])[0].T
        view = camera.inverse(c2w[None])[0].T                                     # the library's stand-in for `.inverse()`
        settings = GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=float((0.5 * fov[0]).tan()), tanfovy=float((0.5 * fov[1]).tan()),
            bg=background, scale_modifier=1.0, viewmatrix=view, projmatrix=view @ proj, sh_degree=degree, campos=c2w[:3, 3],
            prefiltered=False, debug=False)
        scaled_means = means * s
        image, radii = GaussianRasterizer(settings)(
            means3D=scaled_means, means2D=torch.zeros_like(scaled_means, requires_grad=True), shs=sh_upstream,
            colors_precomp=None, opacities=opacities[:, None], cov3D_precomp=(covariances * s ** 2)[:, upper[0], upper[1]])
        assert radii.shape == (means.shape[0],)
        images.append(image)
    return torch.stack(images)


def test_reference_shaped_18_view_loop_through_the_drop_in_module_name():
    from splatter360_b200 import cubemap, synthetic
    from splatter360_b200.decoder import DecoderSplattingCUDA, Gaussians
    dev = torch.device("cuda")
    F, V = 256, 18
    sc = synthetic.pixel_aligned_scene(512, 1024, sh_degree=4, seed=1240, device=dev)
    assert sc.means.shape[0] == 1048576
    poses = synthetic.trajectory(3, seed=6).to(dev)
    ext = cubemap.cube_face_extrinsics(poses).reshape(V, 4, 4)                   # three sets of six cube faces
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev).expand(V, 3, 3)
    near, far = torch.full((V,), 0.5, device=dev), torch.full((V,), 100.0, device=dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    weight = torch.linspace(0.5, 1.5, V * 3 * F * F, device=dev).reshape(V, 3, F, F) / (V * 3 * F * F)

    leaves = [t.clone().requires_grad_() for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]
    loop_img = _stock_view_loop(ext, K, near, far, (F, F), bg, *leaves)
    assert loop_img.shape == (V, 3, F, F)
    (loop_img * weight).sum().backward()
    loop_grads = [t.grad for t in leaves]

    g = Gaussians(*(t[None].clone().requires_grad_() for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)))
    dec = DecoderSplattingCUDA(background_color=tuple(bg.tolist())).to(dev)
    out = dec(g, ext[None], K[None], near[None], far[None], (F, F))
    assert out.color.shape == (1, V, 3, F, F)
    (out.color[0] * weight).sum().backward()

    for v in range(V):
        assert rel_l2(out.color[0, v].detach().cpu().numpy(), loop_img[v].detach().cpu().numpy()) < 1e-6, v
    for a, b, name in zip((g.means, g.covariances, g.harmonics, g.opacities), loop_grads, ("means", "covariances", "harmonics", "opacities")):
        err = rel_l2(a.grad[0].cpu().numpy(), b.cpu().numpy())
        assert err < 2e-5, (name, err)

    # two of the views against the CPU oracle (stock pinhole semantics, the rescaled scene the loop hands over)
    from splatter360_b200 import camera
    for v in (4, 13):
        s = float(1.0 / near[v])
        c2w = ext[v].clone(); c2w[:3, 3] *= s
        cam = camera.pinhole_camera(c2w[None], K[v][None], (near[v] * s)[None], (far[v] * s)[None])
        case = dict(means=(sc.means * s).cpu(), cov6=synthetic.cov3x3_to_cov6(sc.covariances * s ** 2).cpu().contiguous(),
                    opac=sc.opacities.cpu(), shs=sc.harmonics.transpose(-1, -2).contiguous().cpu(), H=F, W=F, mode="pinhole",
                    sh_degree=4, view=cam.view_matrix[0].cpu().contiguous(), proj=cam.full_projection[0].cpu().contiguous(),
                    campos=cam.campos[0].cpu().contiguous(), tanfovx=float(cam.tan_fov_x[0]), tanfovy=float(cam.tan_fov_y[0]),
                    bg=bg.cpu())
        o = run_oracle(case, stages=False)
        err = rel_l2(loop_img[v].detach().cpu().numpy(), o["color"])
        assert err < 1e-4, (v, err)
