"""Drop-in module name: ``from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer`` (/root/reference/src/model/decoder/cuda_splatting.py:5-8) resolves to the
B200-native implementation when this repository is on ``sys.path``."""
from splatter360_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
