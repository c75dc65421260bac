"""GPU parity: CUDA path (through the public GaussianRasterizer / C-ABI) vs the CPU oracle.

Tolerance: north_star asks for <= 1e-4 relative L2 on images and gradients (float32)."""
import numpy as np
import pytest
import torch

from helpers import make_case, rel_l2, run_cuda, run_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("mode,H,W,n", [("pinhole", 96, 128, 3000), ("erp", 64, 128, 3000),
                                         ("pinhole", 50, 70, 500), ("erp", 40, 64, 500)])
def test_forward_backward_parity(mode, H, W, n):
    case = make_case(n, mode, H, W, seed=7)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1))
    o = run_oracle(case, dL=dL)
    c = run_cuda(case, dL=dL)
    assert np.array_equal(o["radii"], c["radii"])
    assert rel_l2(c["color"], o["color"]) < TOL
    for k in ("d_means", "d_cov6", "d_opac", "d_shs", "d_means2D"):
        assert rel_l2(c[k], o[k]) < TOL, k
