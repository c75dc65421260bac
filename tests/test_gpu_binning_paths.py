"""Both binning paths give the same exact lists: the default matrix binning (count matrix over (chunk, tile) -> column
scan -> ranked scatter) and the emit + stable-radix-sort path kept for very large tile counts.  The library reads
S360_FORCE_RADIX_BINNING once per process, so the radix path is exercised in a child pytest process."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_radix_binning_path_still_passes_the_exact_list_tests():
    env = dict(os.environ, S360_FORCE_RADIX_BINNING="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_stages.py",
                        "tests/test_gpu_views.py", "tests/test_gpu_seam_pole.py", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode,H,W,n", [("erp", 1024, 2048, 30000),    # 8192 tiles: the 4-warp scatter instantiation
                                         ("pinhole", 1040, 1040, 20000),  # 4225 tiles, ragged image edge
                                         ("erp", 16, 16, 300),             # one tile
                                         ("pinhole", 200, 300, 2049)])     # chunk boundary: 2048 + 1 Gaussians
def test_matrix_binning_exact_lists_at_the_size_limits(mode, H, W, n):
    import ctypes
    from helpers import make_case, make_settings, run_oracle
    from splatter360_b200 import _lib, rasterizer
    case = make_case(n, mode, H, W, seed=17)
    o = run_oracle(case)
    dev = "cuda"
    s = make_settings(case, dev, tight_bbox=False)
    _, st = rasterizer.forward_raw(s, case["means"].to(dev), case["cov6"].to(dev), case["opac"].to(dev), case["shs"].to(dev), None)
    assert st.num_rendered == o["num_rendered"]
    assert np.array_equal(st.point_list.cpu().numpy().astype(np.uint32)[: st.num_rendered], o["inst_gid"])
    tiles = ((H + 15) // 16) * ((W + 15) // 16)
    rng = torch.zeros(tiles, 2, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().s360_debug_unpack_image(H, W, ctypes.c_void_p(st.image_state.data_ptr()), None, None,
                                                   ctypes.c_void_p(rng.data_ptr()), None))
    torch.cuda.synchronize()
    assert np.array_equal(rng.cpu().numpy().astype(np.uint32), o["tile_ranges"])


def test_many_tiles_fall_back_to_the_radix_path():
    """2048 x 4096 = 32768 tiles is beyond the matrix path's shared-memory counters: emit + radix sort."""
    from helpers import make_case, make_settings, run_oracle, rel_l2
    from splatter360_b200 import rasterizer
    H, W, n = 2048, 4096, 20000
    case = make_case(n, "erp", H, W, seed=23)
    o = run_oracle(case)
    dev = "cuda"
    s = make_settings(case, dev, tight_bbox=False)
    color, st = rasterizer.forward_raw(s, case["means"].to(dev), case["cov6"].to(dev), case["opac"].to(dev), case["shs"].to(dev), None)
    assert st.num_rendered == o["num_rendered"]
    assert np.array_equal(st.point_list.cpu().numpy().astype(np.uint32)[: st.num_rendered], o["inst_gid"])
    assert rel_l2(color.cpu().numpy(), o["color"]) < 1e-4
