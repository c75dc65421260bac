"""Golden arguments of the reference's render_cuda_orthographic (run in the authoring container only).

    python tests/golden/make_golden_ortho.py            # needs /root/reference

ortho_args.npz: what /root/reference/src/model/decoder/cuda_splatting.py:130-220 hands to the rasterizer (captured with a
stub `diff_gaussian_rasterization`) for one batch item -- the reference's
`move_back[2, 3] = -distance_to_near` only works for b = 1 -- at the default fov_degrees = 0.1 (tan(fov/2) ~ 8.7e-4,
camera moved back by ~width/2 / tan) and at fov_degrees = 0.5.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, _load, _ns  # noqa: E402


def inputs(seed=77, G=9, d_sh=25):
    g = torch.Generator().manual_seed(seed)
    ext = torch.eye(4)[None].clone()
    yaw = torch.tensor(0.4)
    ext[0, 0, 0] = yaw.cos(); ext[0, 0, 2] = yaw.sin(); ext[0, 2, 0] = -yaw.sin(); ext[0, 2, 2] = yaw.cos()
    ext[0, :3, 3] = torch.tensor([0.2, -0.1, -1.5])
    means = torch.randn(1, G, 3, generator=g) * 0.6
    A = torch.randn(1, G, 3, 3, generator=g) * 0.08
    cov = A @ A.transpose(-1, -2) + 0.004 * torch.eye(3)
    return dict(extrinsics=ext, width=torch.tensor([3.0]), height=torch.tensor([2.25]), near=torch.tensor([0.05]),
                far=torch.tensor([20.0]), background=torch.rand(1, 3, generator=g), means=means, cov=cov,
                sh=torch.randn(1, G, 3, d_sh, generator=g), op=torch.rand(1, G, generator=g))


def main():
    import types
    for n, p in [("src", "src"), ("src.model", "src/model"), ("src.model.decoder", "src/model/decoder"),
                 ("src.model.encoder", "src/model/encoder"), ("src.model.encoder.costvolume", "src/model/encoder/costvolume"),
                 ("src.geometry", "src/geometry")]:
        _ns(n, os.path.join(REF, p))
    captured = []
    stub = types.ModuleType("diff_gaussian_rasterization")

    class GaussianRasterizationSettings:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class GaussianRasterizer:
        def __init__(self, s):
            self.s = s

        def __call__(self, **kw):
            captured.append((self.s, kw))
            return torch.zeros(3, self.s.image_height, self.s.image_width), torch.zeros(kw["means3D"].shape[0], dtype=torch.int32)

    stub.GaussianRasterizationSettings = GaussianRasterizationSettings
    stub.GaussianRasterizer = GaussianRasterizer
    sys.modules["diff_gaussian_rasterization"] = stub
    _load("src.geometry.projection", f"{REF}/src/geometry/projection.py")
    _load("src.model.encoder.costvolume.conversions", f"{REF}/src/model/encoder/costvolume/conversions.py")
    cs = _load("src.model.decoder.cuda_splatting", f"{REF}/src/model/decoder/cuda_splatting.py")
    d = inputs()
    out = {f"in_{k}": v.numpy() for k, v in d.items()}
    fovs = [0.1, 0.5]
    for fov in fovs:
        dump = {}
        cs.render_cuda_orthographic(d["extrinsics"], d["width"], d["height"], d["near"], d["far"], (48, 64), d["background"],
                                    d["means"], d["cov"], d["sh"], d["op"], fov_degrees=fov, dump=dump)
        s, kw = captured[-1]
        tag = f"call{len(captured) - 1}"
        out[f"{tag}_tanfov"] = np.array([float(s.tanfovx), float(s.tanfovy)], np.float64)
        out[f"{tag}_bg"] = s.bg.numpy(); out[f"{tag}_view"] = s.viewmatrix.numpy(); out[f"{tag}_proj"] = s.projmatrix.numpy()
        out[f"{tag}_campos"] = s.campos.numpy(); out[f"{tag}_sh_degree"] = np.array(s.sh_degree)
        out[f"{tag}_means3D"] = kw["means3D"].numpy(); out[f"{tag}_opacities"] = kw["opacities"].numpy()
        out[f"{tag}_cov6"] = kw["cov3D_precomp"].numpy(); out[f"{tag}_shs"] = kw["shs"].numpy()
        for k, v in dump.items():
            out[f"{tag}_dump_{k}"] = torch.as_tensor(v).numpy()
    out["fov_degrees"] = np.array(fovs)
    np.savez_compressed(os.path.join(OUT, "ortho_args.npz"), **out)
    print("ortho_args.npz written")


if __name__ == "__main__":
    main()
