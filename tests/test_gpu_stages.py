"""Per-stage GPU parity: K1 records, the sorted instance list (exact integer equality),
tile ranges, final_T, n_contrib -- with tight_bbox off so that the instance list is upstream's."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import make_case, make_settings, rel_l2, run_oracle

pytestmark = pytest.mark.gpu


def _unpack(st, P, H, W):
    from splatter360_b200 import _lib
    lib = _lib.load()
    dev = st.geom.device
    xy = torch.zeros(P, 2, device=dev); depth = torch.zeros(P, device=dev); conop = torch.zeros(P, 4, device=dev)
    rgb = torch.zeros(P, 3, device=dev); tiles = torch.zeros(P, dtype=torch.int32, device=dev)
    cl = torch.zeros(P, 3, dtype=torch.uint8, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.s360_debug_unpack_geom(P, p(st.geom), p(xy), p(depth), p(conop), p(rgb), p(tiles), p(cl), None))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    fT = torch.zeros(H, W, device=dev); nc = torch.zeros(H, W, dtype=torch.int32, device=dev)
    rng = torch.zeros(gx * gy, 2, dtype=torch.int32, device=dev)
    _lib.check(lib.s360_debug_unpack_image(H, W, p(st.image_state), p(fT), p(nc), p(rng), None))
    torch.cuda.synchronize()
    return dict(xy=xy.cpu().numpy(), depth=depth.cpu().numpy(), conic_opacity=conop.cpu().numpy(), rgb=rgb.cpu().numpy(),
                tiles_touched=tiles.cpu().numpy().astype(np.uint32), clamped=cl.cpu().numpy(),
                final_T=fT.cpu().numpy(), n_contrib=nc.cpu().numpy().astype(np.uint32),
                tile_ranges=rng.cpu().numpy().astype(np.uint32))


@pytest.mark.parametrize("mode,H,W,n", [("pinhole", 96, 128, 3000), ("erp", 64, 128, 3000), ("erp", 48, 64, 600),
                                         ("pinhole", 130, 70, 5000), ("erp", 256, 512, 20000)])
def test_stage_parity(mode, H, W, n):
    from splatter360_b200 import rasterizer
    case = make_case(n, mode, H, W, seed=3)
    o = run_oracle(case)
    s = make_settings(case, "cuda", tight_bbox=False)
    dev = "cuda"
    color, st = rasterizer.forward_raw(s, case["means"].to(dev), case["cov6"].to(dev), case["opac"].to(dev),
                                       case["shs"].to(dev), None)
    c = _unpack(st, n, H, W)
    vis = o["radii"] > 0
    assert st.num_rendered == o["num_rendered"]
    assert st.num_visible == int(vis.sum())
    assert np.array_equal(st.radii.cpu().numpy(), o["radii"])
    assert np.array_equal(c["tiles_touched"], o["tiles_touched"])
    for k in ("xy", "depth", "conic_opacity", "rgb"):
        assert rel_l2(c[k][vis], o[k][vis]) < 1e-5, k
    assert np.array_equal(c["clamped"][vis], o["clamped"][vis])
    pl = st.point_list.cpu().numpy().astype(np.uint32)[:st.num_rendered]
    assert np.array_equal(pl, o["inst_gid"]), "sorted instance list differs"
    assert np.array_equal(c["tile_ranges"], o["tile_ranges"])
    # last-contributor indices are exact except where an alpha / transmittance threshold decision flips
    # between libm expf and ex2.approx (a handful of pixels in the larger cases)
    mism = float((c["n_contrib"] != o["n_contrib"]).mean())
    assert mism <= (0.0 if n <= 5000 else 2e-3), f"n_contrib mismatch fraction {mism}"
    tol = 1e-5 if n <= 5000 else 1e-4   # a flipped threshold decision moves one pixel by up to ~1/255
    assert rel_l2(c["final_T"], o["final_T"]) < tol
    assert rel_l2(color.cpu().numpy(), o["color"]) < tol
