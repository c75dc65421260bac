"""splatter360_b200 -- B200-native differentiable Gaussian-splat rasterizer (pinhole + native ERP).

Drop-in for the ``diff_gaussian_rasterization`` package that splatter360 imports at
/root/reference/src/model/decoder/cuda_splatting.py:5-8.
"""
__version__ = "0.1.0"
