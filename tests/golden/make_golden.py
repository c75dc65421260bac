"""Generate golden vectors from the REFERENCE's own Python code (run in the authoring container only).

    python tests/golden/make_golden.py            # needs /root/reference

The rasterizer kernels themselves are an absent pip dependency (parity unpinned), but
everything AROUND the kernel call is importable reference code and is pinned here:

  render_args.npz   arguments the reference's render_cuda / render_depth_cuda hand to the rasterizer
                    (/root/reference/src/model/decoder/cuda_splatting.py:47-127, 226-269), captured with a stub
                    `diff_gaussian_rasterization` module
  camera.npz        get_fov / get_projection_matrix (/root/reference/src/geometry/projection.py:233-247;
                    cuda_splatting.py:17-44), depth_to_relative_disparity (conversions.py:17-27)
  erp.npz           hm3d ERP convention: pixel -> direction and point -> pixel
                    (/root/reference/src/geometry/utils360.py:93-104,148-153,193-198,250-263)
  cube2equirec.npz  Cube2Equirec sample grid + one forward (/root/reference/src/geometry/layers.py:41-116)
  covariance.npz    build_covariance (/root/reference/src/model/encoder/common/gaussians.py:8-44)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _ns(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    # bare namespace packages so that the heavy package __init__s (Lightning, dacite, ...) never run
    for n, p in [("src", "src"), ("src.model", "src/model"), ("src.model.decoder", "src/model/decoder"),
                 ("src.model.encoder", "src/model/encoder"), ("src.model.encoder.costvolume", "src/model/encoder/costvolume"),
                 ("src.model.encoder.common", "src/model/encoder/common"), ("src.geometry", "src/geometry")]:
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
    proj = _load("src.geometry.projection", f"{REF}/src/geometry/projection.py")
    _load("src.model.encoder.costvolume.conversions", f"{REF}/src/model/encoder/costvolume/conversions.py")
    cs = _load("src.model.decoder.cuda_splatting", f"{REF}/src/model/decoder/cuda_splatting.py")

    g = torch.Generator().manual_seed(42)
    b, G, d_sh = 2, 7, 25
    yaw = torch.tensor([0.3, -0.7])
    ext = torch.eye(4).repeat(b, 1, 1)
    ext[:, 0, 0] = yaw.cos(); ext[:, 0, 2] = yaw.sin(); ext[:, 2, 0] = -yaw.sin(); ext[:, 2, 2] = yaw.cos()
    ext[:, :3, 3] = torch.randn(b, 3, generator=g) * 0.3
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]]).repeat(b, 1, 1)
    K[1, 0, 0] = 0.8; K[1, 1, 1] = 0.6
    near = torch.tensor([0.1, 0.25]); far = torch.tensor([10.0, 20.0])
    bgc = torch.rand(b, 3, generator=g)
    means = torch.randn(b, G, 3, generator=g) + torch.tensor([0, 0, 3.0])
    A = torch.randn(b, G, 3, 3, generator=g) * 0.1
    cov = A @ A.transpose(-1, -2) + 0.01 * torch.eye(3)
    sh = torch.randn(b, G, 3, d_sh, generator=g)
    op = torch.rand(b, G, generator=g)
    inputs = dict(extrinsics=ext, intrinsics=K, near=near, far=far, background=bgc, means=means, cov=cov, sh=sh, op=op)

    cs.render_cuda(ext, K, near, far, (48, 64), bgc, means, cov, sh, op)
    n_color = len(captured)
    cs.render_depth_cuda(ext, K, near, far, (48, 64), means, cov, op, mode="depth")
    cs.render_depth_cuda(ext, K, near, far, (48, 64), means, cov, op, mode="relative_disparity")
    out = {f"in_{k}": v.numpy() for k, v in inputs.items()}
    for i, (s, kw) in enumerate(captured):
        tag = f"call{i}"
        out[f"{tag}_tanfov"] = np.array([s.tanfovx, s.tanfovy], np.float64)
        out[f"{tag}_hw"] = np.array([s.image_height, s.image_width])
        out[f"{tag}_bg"] = s.bg.numpy(); out[f"{tag}_view"] = s.viewmatrix.numpy(); out[f"{tag}_proj"] = s.projmatrix.numpy()
        out[f"{tag}_campos"] = s.campos.numpy(); out[f"{tag}_sh_degree"] = np.array(s.sh_degree)
        out[f"{tag}_means3D"] = kw["means3D"].numpy(); out[f"{tag}_opacities"] = kw["opacities"].numpy()
        out[f"{tag}_cov6"] = kw["cov3D_precomp"].numpy()
        if kw["shs"] is not None:
            out[f"{tag}_shs"] = kw["shs"].numpy()
        if kw["colors_precomp"] is not None:
            out[f"{tag}_colors"] = kw["colors_precomp"].numpy()
    out["n_calls"] = np.array(len(captured)); out["n_color_calls"] = np.array(n_color)
    np.savez_compressed(os.path.join(OUT, "render_args.npz"), **out)

    fov = proj.get_fov(K)
    pm = cs.get_projection_matrix(near, far, fov[:, 0], fov[:, 1])
    conv = sys.modules["src.model.encoder.costvolume.conversions"]
    depth = torch.rand(2, 5, generator=g) * 9 + 0.2
    rd = conv.depth_to_relative_disparity(depth, near[:, None], far[:, None])
    np.savez_compressed(os.path.join(OUT, "camera.npz"), K=K.numpy(), near=near.numpy(), far=far.numpy(), fov=fov.numpy(),
                        proj=pm.numpy(), depth=depth.numpy(), rel_disp=rd.numpy())

    u360 = _load("src.geometry.utils360", f"{REF}/src/geometry/utils360.py")
    H, W = 32, 64
    ut = u360.Utils({"dataset_name": "hm3d", "batch_size": 1, "height": H, "width": W})
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = torch.stack([xs, ys], -1).reshape(1, -1, 2)
    sph = ut.equi_2_spherical(coords)
    dirs = ut.spherical_2_cartesian(sph)
    pts = torch.randn(1, 200, 3, generator=g) * 2
    sph2 = ut.cartesian_2_spherical(pts)
    pix = ut.spherical_2_equi(sph2)
    np.savez_compressed(os.path.join(OUT, "erp.npz"), H=np.array(H), W=np.array(W), dirs=dirs.reshape(H, W, 3).numpy(),
                        points=pts[0].numpy(), pixels=pix.reshape(-1, 2).numpy())

    layers = _load("src.geometry.layers", f"{REF}/src/geometry/layers.py")
    c2e = layers.Cube2Equirec(8, 16, 32)
    cube = torch.rand(1, 3, 8, 48, generator=g)
    np.savez_compressed(os.path.join(OUT, "cube2equirec.npz"), cube=cube.numpy(), erp=c2e(cube).detach().numpy(),
                        grid=c2e.sample_grid.detach().numpy())

    gs = _load("src.model.encoder.common.gaussians", f"{REF}/src/model/encoder/common/gaussians.py")
    scale = torch.rand(6, 3, generator=g) + 0.1
    quat = torch.randn(6, 4, generator=g)
    np.savez_compressed(os.path.join(OUT, "covariance.npz"), scale=scale.numpy(), quat_xyzw=quat.numpy(),
                        cov=gs.build_covariance(scale, quat).numpy())
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
