"""Golden vectors of the reference's GaussianAdapterERP (run in the authoring container only).

    python tests/golden/make_golden_adapter.py            # needs /root/reference

adapter.npz: inputs, outputs and input-gradients of /root/reference/src/model/encoder/common/gaussian_adapter_erp.py:49-119
run UNMODIFIED (scale activation, quaternion normalisation, build_covariance gaussians.py:8-44, rotation into the world
frame, ERP unprojection sphere_projection.py:6-87 with the hm3d convention of utils360.py, SH mask).  The one thing the
reference cannot supply here is ``e3nn`` (sh_rotation.py:4): ``e3nn.o3`` is stubbed with oracle/e3nn_wigner.py, the
restatement of e3nn's published matrix_to_angles / wigner_D -- so means, covariances and opacities are pinned by the
reference's own code, the rotated harmonics by the reference's code ON TOP of that restatement (marked unpinned).
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from make_golden import OUT, REF, _load, _ns  # noqa: E402


def load_reference_adapter():
    from oracle import e3nn_wigner
    e3nn = types.ModuleType("e3nn"); o3 = types.ModuleType("e3nn.o3")
    o3.matrix_to_angles, o3.wigner_D = e3nn_wigner.matrix_to_angles, e3nn_wigner.wigner_D
    e3nn.o3 = o3
    sys.modules["e3nn"], sys.modules["e3nn.o3"] = e3nn, o3
    for n, p in [("src", "src"), ("src.model", "src/model"), ("src.model.encoder", "src/model/encoder"),
                 ("src.model.encoder.common", "src/model/encoder/common"), ("src.geometry", "src/geometry"), ("src.misc", "src/misc")]:
        _ns(n, os.path.join(REF, p))
    _load("src.geometry.projection", f"{REF}/src/geometry/projection.py")
    _load("src.geometry.utils360", f"{REF}/src/geometry/utils360.py")
    sp = _load("src.geometry.sphere_projection", f"{REF}/src/geometry/sphere_projection.py")
    # the reference writes the depth in place into an einops-expanded tensor, which only works after its .to(cuda) copy:
    # on this CPU-only container the harness makes einops.repeat return real memory (values unchanged)
    import einops
    sp.repeat = lambda *a, **k: einops.repeat(*a, **k).clone()
    _load("src.misc.sh_rotation", f"{REF}/src/misc/sh_rotation.py")
    _load("src.model.encoder.common.gaussians", f"{REF}/src/model/encoder/common/gaussians.py")
    _load("src.model.encoder.common.gaussian_adapter", f"{REF}/src/model/encoder/common/gaussian_adapter.py")
    return _load("src.model.encoder.common.gaussian_adapter_erp", f"{REF}/src/model/encoder/common/gaussian_adapter_erp.py")


def inputs(seed=5, b=1, v=2, h=8, w=16, sh_degree=4):
    g = torch.Generator().manual_seed(seed)
    d_sh = (sh_degree + 1) ** 2
    r = h * w
    ext = torch.eye(4).repeat(b, v, 1, 1)
    for i in range(b):
        for j in range(v):
            q = torch.randn(4, generator=g); q = q / q.norm()
            x, y, z, s = q
            ext[i, j, :3, :3] = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * s), 2 * (x * z + y * s)],
                                              [2 * (x * y + z * s), 1 - 2 * (x * x + z * z), 2 * (y * z - x * s)],
                                              [2 * (x * z - y * s), 2 * (y * z + x * s), 1 - 2 * (x * x + y * y)]])
            ext[i, j, :3, 3] = torch.randn(3, generator=g) * 0.5
    return dict(extrinsics=ext, depths=0.5 + 5.0 * torch.rand(b, v, r, 1, 1, generator=g),
                opacities=torch.rand(b, v, r, 1, 1, generator=g), raw=torch.randn(b, v, r, 1, 1, 7 + 3 * d_sh, generator=g))


def main():
    mod = load_reference_adapter()
    cfg = mod.GaussianAdapterERPCfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, sh_degree=4)   # config/model/encoder/costvolume.yaml
    adapter = mod.GaussianAdapterERP(cfg)
    d = inputs()
    h, w = 8, 16
    depths = d["depths"].clone().requires_grad_()
    raw = d["raw"].clone().requires_grad_()
    ext = d["extrinsics"][:, :, None, None, None]
    out = adapter.forward("hm3d", ext, depths, d["opacities"], raw, (h, w))
    g = torch.Generator().manual_seed(9)
    cm, cc, ch = (torch.randn(t.shape, generator=g) for t in (out.means, out.covariances, out.harmonics))
    ((out.means * cm).sum() + (out.covariances * cc).sum() + (out.harmonics * ch).sum()).backward()
    np.savez_compressed(
        os.path.join(OUT, "adapter.npz"), h=np.array(h), w=np.array(w), scale_min=np.array(0.5), scale_max=np.array(15.0),
        extrinsics=d["extrinsics"].numpy(), depths=d["depths"].numpy(), opacities=d["opacities"].numpy(), raw=d["raw"].numpy(),
        means=out.means.detach().numpy(), covariances=out.covariances.detach().numpy(), harmonics=out.harmonics.detach().numpy(),
        out_opacities=out.opacities.numpy(), scales=out.scales.detach().numpy(), rotations=out.rotations.detach().numpy(),
        cot_means=cm.numpy(), cot_cov=cc.numpy(), cot_sh=ch.numpy(), d_depths=depths.grad.numpy(), d_raw=raw.grad.numpy(),
        sh_mask=adapter.sh_mask.numpy())
    print("adapter.npz written", {k: tuple(getattr(out, k).shape) for k in ("means", "covariances", "harmonics", "opacities")})


if __name__ == "__main__":
    main()
