"""Golden vector for the reference's ERP depth panorama (run in the authoring container only; needs /root/reference):

    six z-depth cube faces (dataset order [U B L F R D])
      -> change_order_batch                      /root/reference/src/model/model_wrapper_erp.py:147-158 (function source exec'd;
                                                  the module itself needs Lightning)
      -> depth_to_distance_map_batch             /root/reference/src/geometry/z_depth_to_distance.py:4-34
      -> strip [v,1,h,6w] -> Cube2Equirec        model_wrapper_erp.py:447-463, src/geometry/layers.py:41-116

writes tests/golden/depth_panorama.npz.  The chain is what `Cube2Equirec.from_faces(..., depth_to_distance=...)` fuses.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from einops import rearrange, repeat

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    src = open(f"{REF}/src/model/model_wrapper_erp.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "change_order_batch"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "model_wrapper_erp.py", "exec"), ns)
    change_order_batch = ns["change_order_batch"]
    z2d = _load("z_depth_to_distance", f"{REF}/src/geometry/z_depth_to_distance.py")
    pkg = types.ModuleType("src"); pkg.__path__ = [f"{REF}/src"]; sys.modules["src"] = pkg
    geo = types.ModuleType("src.geometry"); geo.__path__ = [f"{REF}/src/geometry"]; sys.modules["src.geometry"] = geo
    layers = _load("src.geometry.layers", f"{REF}/src/geometry/layers.py")

    g = torch.Generator().manual_seed(77)
    v, f, H, W = 2, 16, 32, 64
    faces = 0.5 + 4 * torch.rand(v, 6, f, f, generator=g)            # z-depth, dataset order
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])   # normalised 90-degree face intrinsics
    reordered = change_order_batch(faces.clone())
    reordered = rearrange(reordered, "v cubes h w -> (v cubes) h w")
    intr = K[None].repeat(v * 6, 1, 1)
    fx, fy, cx, cy = intr[:, 0, 0] * f, intr[:, 1, 1] * f, intr[:, 0, 2] * f, intr[:, 1, 2] * f
    fxfycxcy = repeat(torch.stack([fx, fy, cx, cy], dim=1), "vc r -> vc r h w", h=f, w=f)
    dist = z2d.depth_to_distance_map_batch(reordered, fxfycxcy)
    strip = rearrange(dist, "(v cubes) h w -> v () h (cubes w)", v=v, cubes=6)
    pano = layers.Cube2Equirec(f, H, W)(strip).squeeze(1)
    np.savez_compressed(os.path.join(OUT, "depth_panorama.npz"), faces=faces.numpy(), pano=pano.detach().numpy(),
                        fxfycxcy=np.array([float(fx[0]), float(fy[0]), float(cx[0]), float(cy[0])], dtype=np.float32),
                        face_w=np.array(f), H=np.array(H), W=np.array(W))
    print("written", os.path.join(OUT, "depth_panorama.npz"), pano.shape)


if __name__ == "__main__":
    main()
